"""CPU: pins the oracle's restated optimiser step against the reference's UNMODIFIED `MyHandler._update_disc` /
`_update_gen` (model/model_handler.py:349-498) run here on the host (its `.cuda()` calls shimmed to identities,
oracle/ref_harness.py), and checks the four-import swap machinery the GPU drop-in test relies on.  Skipped when neither
/root/reference nor the staged copy oracle/_ref is present."""
import numpy as np
import pytest
import torch

from oracle import advmil_oracle as O
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not available (oracle/build_ref.py)")


def _handler(tmp_path, impl, **over):
    from oracle import ref_harness as H
    paths = H.write_synthetic_dataset(str(tmp_path / "data"), 8, 320, seed=7)
    return H, H.make_handler(H.load_cfg(**paths, **over), impl, "cpu")


def test_oracle_step_equals_unmodified_handler_step(tmp_path):
    """Two optimiser steps (D update + G update, Adam, L1, weight-decay rule) over 6 ragged bags with mixed labels: the
    oracle trainer and the reference's own handler methods agree on losses, predictions and every parameter."""
    H, h = _handler(tmp_path, "reference")
    H.set_dropout(h.netG, 0.0)
    H.set_dropout(h.netD, 0.0)
    sdG = {k: v.detach().clone() for k, v in h.netG.state_dict().items()}
    sdD = {k: v.detach().clone() for k, v in h.netD.state_dict().items()}
    tr = O.CpuTrainer(sdG, sdD)
    Ns = [320, 160, 480, 16, 640, 336]
    bags = [O.synth_bag(n, 30 + i) for i, n in enumerate(Ns)]
    ts, es = torch.tensor([0.3, 0.8, 0.55, 0.1, 0.9, 0.42]), torch.tensor([1.0, 0.0, 1.0, 0.0, 0.0, 1.0])
    xs = [[b.unsqueeze(0), torch.Tensor([0]).unsqueeze(0)] for b in bags]
    ys = [torch.stack([t, e]).reshape(1, 2) for t, e in zip(ts, es)]
    for step in range(2):
        # the handler draws its noise from the CPU generator (utils/func.py:154-164): replay the same stream for the oracle
        state = torch.get_rng_state()
        nzd = [torch.rand(1, 192) for _ in Ns]
        nzg = [torch.rand(1, 192) for _ in Ns]
        torch.set_rng_state(state)
        with H.device_context("cpu"):
            preds, fakes, log = H.handler_step(h, 16 * (step + 1), xs, ys)
        o = tr.step(bags, ts, es, [True] * len(Ns), nzd, nzg)
        L = H.parse_losses(log)
        assert abs(L["dis_loss"][0] - o["dis_loss"]) < 2e-6 and abs(L["gen_loss"][0] - o["gen_loss"]) < 2e-6
        assert abs(L["t_reg_loss"][0] - o["t_reg"]) < 2e-6 and abs(L["gen_totol_loss"][0] - o["total"]) < 2e-6
        np.testing.assert_allclose(torch.cat(preds).detach().reshape(-1).numpy(), o["pred_d"].reshape(-1).numpy(), rtol=0, atol=1e-6)
        np.testing.assert_allclose(torch.cat(fakes).reshape(-1).numpy(), o["fake_d"].reshape(-1).numpy(), rtol=0, atol=2e-6)
    for k, v in h.netG.state_dict().items():
        np.testing.assert_allclose(v.numpy(), tr.sdG[k].detach().numpy(), rtol=0, atol=2e-6, err_msg=k)
    for k, v in h.netD.state_dict().items():
        np.testing.assert_allclose(v.numpy(), tr.sdD[k].detach().numpy(), rtol=0, atol=2e-6, err_msg=k)


@pytest.mark.parametrize("mode", ["abmil", "patch", "cluster"])
def test_import_swap_builds_identically_initialised_networks(tmp_path, mode):
    """`MyHandler.__init__` (model_handler.py:74-91) around the advmil_b200 modules consumes the RNG exactly like the
    reference's: same state_dict keys, bit-identical initial values (seed_everything(42) + init_weights on G)."""
    H, h1 = _handler(tmp_path, "reference", bcb_mode=mode)
    _, h2 = _handler(tmp_path, "advmil_b200", bcb_mode=mode)
    assert type(h2.netG).__module__.startswith("advmil_b200.") and type(h1.netG).__module__ == "model.GANSurv"
    for a, b in ((h1.netG, h2.netG), (h1.netD, h2.netD)):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        for k in sa:
            assert torch.equal(sa[k], sb[k]), k
    # the optimiser's no-decay rule sees the same parameter groups (optim/optim_factory.py:25-37)
    g1 = [[tuple(p.shape) for p in g["params"]] for g in h1.optimizerG.param_groups]
    g2 = [[tuple(p.shape) for p in g["params"]] for g in h2.optimizerG.param_groups]
    assert g1 == g2
