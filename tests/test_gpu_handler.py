"""GPU: the reference's UNMODIFIED `MyHandler` (model/model_handler.py) driving the advmil_b200 modules.

`oracle/ref_harness.make_handler(cfg, "advmil_b200")` executes the reference's own model_handler.py with its four model
imports (:13-16) resolved to advmil_b200 -- nothing else of the handler changes -- and `make_handler(cfg, "reference")`
is the stock handler around the reference's modules on the same GPU.  Both run `_train_each_epoch` (:301-347) over the
same synthetic per-slide `.pt` files through the reference's own `WSIPatch` dataset and a DataLoader, and
`test_model` (:598-643).  The dropout generators differ by design (Philox vs counter hash), so the epoch comparison
runs with every dropout probability set to 0 on both module trees; noise comes from the CPU stream and is identical."""
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle import ref_import
from tests.util import assert_close

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_import.available(), reason="reference tree not staged (oracle/build_ref.py)")]

ZERO_GRAD = ("pool.fc2.bias", "attention_c.bias")
ROWS = [1600, 320, 4096, 640, 16, 2048, 960, 1280] * 4       # 32 patients = two 16-bag optimiser steps


def _setup(tmp_path, mode="abmil", **over):
    from oracle import ref_harness as H
    paths = H.write_synthetic_dataset(str(tmp_path / "data"), len(ROWS), ROWS, seed=11)
    cfg = H.load_cfg(**paths, bcb_mode=mode, **over)
    return H, cfg


def _loader(H, h, cfg):
    from dataset.utils import prepare_dataset          # the reference's own dataset code (dataset/PatchWSI.py)
    from torch.utils.data import DataLoader
    from utils.io import read_datasplit_npz
    pids_train, pids_val, _ = read_datasplit_npz(cfg["data_split_path"].format(0))
    with contextlib.redirect_stdout(io.StringIO()):
        ds = prepare_dataset(pids_train, cfg)
    h.patient_id.update({"label_visible": pids_train + pids_val, "train": ds.pids})
    return DataLoader(ds, batch_size=1, shuffle=False, num_workers=0)


def _epoch(H, impl, cfg, p_drop=None, seed=5):
    h = H.make_handler(cfg, impl, "cuda")
    if p_drop is not None:
        H.set_dropout(h.netG, p_drop)
        H.set_dropout(h.netD, p_drop)
    dl = _loader(H, h, cfg)
    torch.manual_seed(seed)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        cltor = h._train_each_epoch(dl, "train")
    torch.cuda.synchronize()
    return h, cltor, H.parse_losses(sink.getvalue())


@pytest.fixture()
def exact_reference_arithmetic():
    """cuDNN's conv (D's 1x1 conv, backbone_utils.py:143) defaults to TF32 on this GPU; parity wants the reference's fp32."""
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved


def test_unmodified_handler_epoch_dropin_equals_reference(tmp_path, exact_reference_arithmetic):
    import advmil_b200
    H, cfg = _setup(tmp_path)
    advmil_b200.set_precision("fp32")
    init = {k: v.clone() for k, v in H.make_handler(cfg, "reference", "cuda").netG.state_dict().items()}
    href, cref, Lref = _epoch(H, "reference", cfg, p_drop=0.0)
    hnew, cnew, Lnew = _epoch(H, "advmil_b200", cfg, p_drop=0.0)
    assert type(hnew.netG).__module__.startswith("advmil_b200.")
    assert len(Lref["dis_loss"]) == 2 and len(Lnew["dis_loss"]) == 2
    for k in Lref:                                     # the handler prints six decimals
        np.testing.assert_allclose(Lnew[k], Lref[k], rtol=0, atol=3e-5, err_msg=k)
    assert_close(cnew["y_hat"], cref["y_hat"], 1e-5, "epoch y_hat")
    assert_close(cnew["f_fake"], cref["f_fake"], 1e-5, "epoch f_fake", atol=2e-6)
    assert torch.equal(cnew["y"], cref["y"])
    # two Adam steps per network: a parameter moves by <= 2 * lr = 1.6e-4; agreement to a small fraction of that.  The two
    # biases whose gradient is mathematically zero (softmax shift invariance) move by +-lr on rounding noise of either
    # implementation under Adam's normalisation: only their magnitude is bounded.
    moved = 0.0
    for new_net, ref_net in ((hnew.netG, href.netG), (hnew.netD, href.netD)):
        for (k, a), (_, b) in zip(new_net.state_dict().items(), ref_net.state_dict().items()):
            if k.endswith(ZERO_GRAD):
                assert float((a - b).abs().max()) <= 2.1 * 2 * 8e-5, k
                continue
            assert float((a - b).abs().max()) <= 0.05 * 2 * 8e-5, (k, float((a - b).abs().max()))
            moved = max(moved, float((b - init[k]).abs().max()) if k in init else 0.0)
    assert moved > 8e-5


def test_unmodified_handler_test_model_dropin_equals_reference(tmp_path, exact_reference_arithmetic):
    """`MyHandler.test_model` (:598-643): 1 + 30 generator forwards per bag, D(x, y_hat), lower median."""
    import advmil_b200
    H, cfg = _setup(tmp_path)
    advmil_b200.set_precision("fp32")
    res = {}
    for impl in ("reference", "advmil_b200"):
        h = H.make_handler(cfg, impl, "cuda")
        dl = _loader(H, h, cfg)
        torch.manual_seed(99)
        res[impl] = h.test_model(h.netG, h.netD, "abmil", dl, times_test_sample=30)
    a, b = res["advmil_b200"], res["reference"]
    assert tuple(a["dist_y_hat"].shape) == (len(ROWS), 30, 1)
    for k in ("y_hat", "dist_y_hat", "avg_y_hat"):
        assert_close(a[k], b[k], 1e-5, k)
    assert_close(a["f_fake"], b["f_fake"], 1e-5, "f_fake", atol=2e-6)
    assert torch.equal(a["idx"], b["idx"]) and torch.equal(a["y"], b["y"])


@pytest.mark.parametrize("precision", ["bf16", "tf32x3"])
def test_unmodified_handler_epoch_in_the_fast_modes(tmp_path, exact_reference_arithmetic, precision):
    """The same epoch with the drop-in modules in the bf16 storage mode (2e-2, north_star) and in the split-tf32
    tensor-core mode (fp32-grade): losses and predictions against the stock handler on the same GPU."""
    import advmil_b200
    H, cfg = _setup(tmp_path)
    href, cref, Lref = _epoch(H, "reference", cfg, p_drop=0.0)
    advmil_b200.set_precision(precision)
    try:
        hnew, cnew, Lnew = _epoch(H, "advmil_b200", cfg, p_drop=0.0)
    finally:
        advmil_b200.set_precision("fp32")
    tol = 2e-2 if precision == "bf16" else 1e-4
    for k in Lref:
        np.testing.assert_allclose(Lnew[k], Lref[k], rtol=tol, atol=tol * 0.1, err_msg=k)
    assert_close(cnew["y_hat"], cref["y_hat"], tol, "epoch y_hat")


def test_unmodified_handler_with_dropout_trains(tmp_path):
    """Stock configuration (dropout on): the run cannot be compared element-wise across dropout generators; the eval-mode
    quantities of the first optimiser step (G.eval predictions, model_handler.py:355-356,383-391) still agree exactly,
    and the epoch's losses stay finite and close to the stock handler's."""
    import advmil_b200
    H, cfg = _setup(tmp_path)
    advmil_b200.set_precision("fp32")
    _, cref, Lref = _epoch(H, "reference", cfg)
    _, cnew, Lnew = _epoch(H, "advmil_b200", cfg)
    assert_close(cnew["y_hat"][:16], cref["y_hat"][:16], 1e-5, "first-step eval predictions")
    for k in Lref:
        assert np.all(np.isfinite(Lnew[k]))
        np.testing.assert_allclose(Lnew[k], Lref[k], rtol=0.2, atol=0.1, err_msg=k)
