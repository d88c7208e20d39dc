"""GPU parity of the fused adversarial step (AdvStep) against fixtures produced by the live reference's modules,
losses and optimisers (tests/golden/step_*.npz, oracle/make_golden.py:case_step) — D step + G step + both Adam
updates, train-mode dropout with injected masks."""
import numpy as np
import pytest
import torch

from oracle import advmil_oracle as O
from tests.util import assert_close, build_D, build_G, d_masks, esat_masks, g_masks, golden, grad_floor, sub

pytestmark = pytest.mark.gpu
RTOL = 1e-5
ZERO_GRAD = ("pool.fc2.bias", "attention_c.bias")   # mathematically-zero gradients: Adam amplifies their rounding noise


def _cat_masks(per_bag, keys):
    return {k: torch.cat([m[k] for m in per_bag], dim=0).to(torch.uint8).contiguous().cuda() for k in keys}


@pytest.mark.parametrize("name", ["step_small", "step_full"])
def test_fused_step_vs_reference_golden(name):
    from advmil_b200 import ops
    from advmil_b200.step import AdvStep
    g = golden(name)
    C, h, o, d, seed, n_steps = [int(v) for v in g["cfg"][:6]]
    Ns = [int(v) for v in g["cfg"][6:]]
    B = len(Ns)
    sdG = O.synth_state_dict(O.G_SHAPES(C, h, o), seed)
    sdD = O.synth_state_dict(O.D_SHAPES(C, d, (64, 128) if d == 128 else (d // 2, d)), seed + 50)
    G, D = build_G((C, h, o)), build_D(C, d)
    G.load_state_dict(sdG)
    D.load_state_dict(sdD)
    eng = AdvStep(G, D)
    bags = ops.PackedBags.from_list([O.synth_bag(n, seed + i, C).cuda() for i, n in enumerate(Ns)])
    t, e = torch.tensor(g["t"]).cuda(), torch.tensor(g["e"]).cuda()
    vis = torch.tensor(g["visible"].astype(np.uint8)).cuda()
    for step in range(n_steps):
        rng = np.random.default_rng(seed + 100 * step)
        nzD = torch.tensor(np.concatenate([rng.uniform(size=(1, o // 2)) for _ in range(B)]), dtype=torch.float32).cuda()
        nzG = torch.tensor(np.concatenate([rng.uniform(size=(1, o // 2)) for _ in range(B)]), dtype=torch.float32).cuda()
        mr = _cat_masks([d_masks(Ns[i] // 16, d, seed + 1000 * step + 10 * i) for i in range(B)], ["fc1", "ga", "gs", "fc2"])
        mf = _cat_masks([d_masks(Ns[i] // 16, d, seed + 1000 * step + 10 * i + 5) for i in range(B)], ["fc1", "ga", "gs", "fc2"])
        mg = _cat_masks([g_masks(Ns[i], h, o, seed + 2000 * step + 10 * i) for i in range(B)], ["h", "a", "b", "rho", "mlp0"])
        out = eng.step(bags, t, e, vis, noise_d=nzD, noise_g=nzG, masks_d_real=mr, masks_d_fake=mf, masks_g=mg)
        L = eng.loss_dict(out)
        assert_close(out["pred_d"].cpu(), g[f"pred_d{step}"], RTOL, f"pred_d step {step}")
        assert_close(out["pred_g"].cpu(), g[f"pred_g{step}"], RTOL, f"pred_g step {step}")
        assert_close(out["f_fake_d"].cpu(), g[f"fake_d{step}"], RTOL, f"fake_d step {step}", atol_scale=1e-1)
        assert_close(out["f_fake_g"].cpu(), g[f"fake_g{step}"], RTOL, f"fake_g step {step}", atol_scale=1e-1)
        assert abs(L["dis_loss"] - float(g[f"dis_loss{step}"])) < 2e-5
        assert abs(L["gen_loss"] - float(g[f"gen_loss{step}"])) < 2e-5
        assert abs(L["t_reg_loss"] - float(g[f"t_reg{step}"])) < 2e-5
        assert abs(L["gen_total_loss"] - float(g[f"total{step}"])) < 2e-5
        if step == 0:
            dnames = [k for k, _ in D.named_parameters()]
            # real (-) and fake (+) pair contributions of size ~(1/B)*O(1) cancel in the head's bias gradients: the fp32
            # rounding floor is one ulp of those terms, 2^-23 / B, whatever the size of the (much smaller) sum
            floor = max(grad_floor([g["dgrad." + k] for k in dnames]), 2.0 ** -23 / B)
            grads = dict(zip(dnames, [eng.dgrads[i] for i in _order(D, eng.dparams)]))
            for k in dnames:
                if k.endswith(ZERO_GRAD):
                    continue
                assert_close(sub(grads[k]), g["dgrad." + k], RTOL, "D grad " + k, atol=floor)
            gnames = [k for k, _ in G.named_parameters()]
            floor = grad_floor([g["ggrad." + k] for k in gnames])
            ggr = dict(zip(gnames, [eng.ggrads[i] for i in _order(G, eng.gparams)]))
            for k in gnames:
                if k.endswith(ZERO_GRAD):
                    continue
                full = ggr[k].cpu() + 1e-5 * torch.sign(sdG[k])   # loss_reg_l1 (loss/utils.py:6-14) is folded into Adam
                assert_close(sub(full), g["ggrad." + k], RTOL, "G grad " + k, atol=floor)
    # parameters after the Adam updates (lr 8e-5): compare the UPDATE, |delta - delta_ref| <= 1e-3 * lr-scale
    for k, p in G.named_parameters():
        if k.endswith(ZERO_GRAD):
            continue
        assert_close(sub(p), g["gparam." + k], RTOL, "G param " + k, atol=8e-5 * n_steps * 2e-2)
    for k, p in D.named_parameters():
        if k.endswith(ZERO_GRAD):
            continue
        assert_close(sub(p), g["dparam." + k], RTOL, "D param " + k, atol=8e-5 * n_steps * 2e-2)


def _order(mod, tensors):
    """indices into `tensors` (engine order, Nones dropped) following mod.named_parameters() order"""
    pos = {id(t): i for i, t in enumerate(tensors) if t is not None}
    return [pos[id(p)] for _, p in mod.named_parameters()]


def test_sample_inference_matches_oracle_and_lower_median():
    from advmil_b200 import ops
    from advmil_b200.step import sample_inference
    dims = (1024, 384, 384)
    sdG, sdD = O.synth_state_dict(O.G_SHAPES(*dims), 3), O.synth_state_dict(O.D_SHAPES(), 4)
    G, D = build_G(dims).eval(), build_D().eval()
    G.load_state_dict(sdG)
    D.load_state_dict(sdD)
    Ns = [320, 640, 160]
    xs = [O.synth_bag(n, 70 + i) for i, n in enumerate(Ns)]
    torch.manual_seed(99)
    res = sample_inference(G, D, ops.PackedBags.from_list([x.cuda() for x in xs]), times_test_sample=30)
    # replay the same CPU RNG stream through the oracle: first draw -> y_hat, then 30 draws -> distribution
    torch.manual_seed(99)
    first = torch.rand(len(Ns), 192)
    draws = [torch.rand(len(Ns), 192) for _ in range(30)]
    for b, x in enumerate(xs):
        y = O.generator_forward(sdG, x, [None, first[b:b + 1]], (0, 1))["pred"]
        assert_close(res["y_hat"][b].cpu(), y.reshape(-1), RTOL, f"y_hat {b}")
        f = O.prjdisc_forward(sdD, x, y)["out"]
        assert_close(res["f_fake"][b].cpu(), f.reshape(-1), RTOL, f"f_fake {b}", atol_scale=1e-1)
        dist = O.sample_times(sdG, x, [d_[b:b + 1] for d_ in draws])
        assert_close(res["dist_y_hat"][b, :, 0].cpu(), dist, RTOL, f"dist {b}")
        med = O.lower_median(dist)
        assert abs(float(res["avg_y_hat"][b]) - float(med)) < 1e-6


@pytest.mark.parametrize("kind", ["patch", "cluster"])
def test_sample_inference_for_the_other_backbones(kind):
    """MyHandler.test_model (model_handler.py:598-643) works for any backbone: sample_inference for the ESAT (`patch`) and
    DeepAttMISL (`cluster`) generators from ONE backbone pass equals the reference's 1 + S per-bag generator forwards
    replayed on the same CPU noise stream through the oracle; lower median; D(x, y_hat)."""
    from advmil_b200 import ops
    from advmil_b200.step import sample_inference
    dims, S = (1024, 384, 384), 7
    shapes = O.G_ESAT_SHAPES() if kind == "patch" else O.G_CLUSTER_SHAPES(dims[0], dims[1])
    sdG, sdD = O.synth_state_dict(shapes, 31), O.synth_state_dict(O.D_SHAPES(), 32)
    G, D = build_G(dims, mode=kind).eval(), build_D().eval()
    G.load_state_dict(sdG)
    D.load_state_dict(sdD)
    Ns = [320, 640, 160]
    xs = [O.synth_bag(n, 170 + i) for i, n in enumerate(Ns)]
    rng = np.random.default_rng(33)
    cids = [torch.tensor(rng.integers(0, 8, size=n), dtype=torch.float32) for n in Ns]
    ext = torch.cat(cids).cuda() if kind == "cluster" else None
    torch.manual_seed(77)
    res = sample_inference(G, D, ops.PackedBags.from_list([x.cuda() for x in xs]), times_test_sample=S, ext=ext)
    torch.manual_seed(77)
    first = torch.rand(len(Ns), 192)
    draws = [torch.rand(len(Ns), 192) for _ in range(S)]
    for b, x in enumerate(xs):
        kw = {"backbone": kind}
        if kind == "cluster":
            kw["cluster_id"] = cids[b]
        y = O.generator_forward(sdG, x, [None, first[b:b + 1]], (0, 1), None, **kw)["pred"]
        assert_close(res["y_hat"][b].cpu(), y.reshape(-1), RTOL, f"y_hat {b}")
        f = O.prjdisc_forward(sdD, x, y)["out"]
        assert_close(res["f_fake"][b].cpu(), f.reshape(-1), RTOL, f"f_fake {b}", atol_scale=1e-1)
        dist = torch.stack([O.generator_forward(sdG, x, [None, d_[b:b + 1]], (0, 1), None, **kw)["pred"].reshape(()) for d_ in draws])
        assert_close(res["dist_y_hat"][b, :, 0].cpu(), dist, RTOL, f"dist {b}")
        assert abs(float(res["avg_y_hat"][b]) - float(O.lower_median(dist))) < 1e-6


def test_device_cindex_matches_reference_golden_and_oracle_counts():
    """eval/cindex.py on the device: the value of the reference's concordance_index on 447 patients with tied times and
    tied predictions (tests/golden/misc.npz, written by the live reference) is reproduced exactly (integer pair counts,
    one double division); random cases against the oracle's pair loop."""
    from advmil_b200.eval.cindex import concordance_counts, concordance_index
    g = golden("misc")
    y_true = np.stack([g["ci_t"], g["ci_e"]], axis=1)
    ci = concordance_index(y_true, g["ci_pred"].reshape(-1, 1))
    assert ci == float(g["ci"]), (ci, float(g["ci"]))
    rng = np.random.default_rng(3)
    for n in (2, 17, 300, 1500):
        t = np.round(rng.uniform(size=n), 2).astype(np.float32)
        e = (rng.uniform(size=n) < 0.4).astype(np.float32)
        e[0] = 1.0
        p = np.round(rng.uniform(size=n), 2).astype(np.float32)
        c = concordance_counts(t, e, p)
        if c["comparable"] == 0:
            continue
        want = O.concordance_index(t, e, p)
        got = (c["concordant"] + 0.5 * c["tied_risk"]) / c["comparable"]
        assert got == want, (n, got, want)
        assert c["concordant"] + c["tied_risk"] + c["discordant"] == c["comparable"]
    with pytest.raises(ValueError):
        concordance_index(np.array([[0.5, 0.0], [0.7, 0.0]], dtype=np.float32), np.array([[0.1], [0.2]], dtype=np.float32))


def test_packed_file_feeder_end_to_end(tmp_path):
    """dataset/packed_file.py + DeviceFeeder + AdvStep (the e2e path of bench.py): bags written once as a bf16 blob, steps
    assembled from the mmap into pinned buffers, copied asynchronously (double buffered), consumed by the fused step.  The
    features that arrive on the device are exactly the stored ones; the step's outputs equal those of the same bags handed
    over directly."""
    from advmil_b200 import ops
    from advmil_b200.dataset.packed import DeviceFeeder, group_steps
    from advmil_b200.dataset.packed_file import PackedFile, write_packed
    from advmil_b200.step import AdvStep
    import advmil_b200.step as step_mod
    g = torch.Generator().manual_seed(3)
    lens = [64, 320, 16, 160, 96, 48, 640, 32]
    bags = [torch.randn(n, 1024, generator=g) for n in lens]
    labels = [(0.1 + 0.1 * i, float(i % 2 == 0)) for i in range(len(lens))]
    path = str(tmp_path / "train.advmil")
    write_packed(path, iter(bags), labels, dtype=torch.bfloat16)
    pf = PackedFile(path)
    groups = group_steps(len(pf), 4)
    assert groups == [[0, 1, 2, 3], [4, 5, 6, 7]]
    sdG, sdD = O.synth_state_dict(O.G_SHAPES(), 5), O.synth_state_dict(O.D_SHAPES(), 6)

    def run(feed):
        seeds = iter(range(1000, 1100))
        step_mod.next_dropout_seed, saved = (lambda: next(seeds)), step_mod.next_dropout_seed
        try:
            G, D = build_G(), build_D()
            G.load_state_dict(sdG)
            D.load_state_dict(sdD)
            eng = AdvStep(G, D, precision="bf16")
            torch.manual_seed(11)
            outs = []
            for bg, t, e, vis in feed():
                outs.append(eng.step(bg, t, e, vis)["pred_g"].clone())
            torch.cuda.synchronize()
            return outs, [p.detach().clone() for p in G.parameters()]
        finally:
            step_mod.next_dropout_seed = saved

    def via_feeder():
        for s, idx in zip(DeviceFeeder((pf.step(idx) for idx in groups), device="cuda"), groups):
            want = torch.cat([bags[i] for i in idx]).to(torch.bfloat16)
            assert s.bags.x.dtype == torch.bfloat16 and torch.equal(s.bags.x.cpu(), want)
            assert s.bags.offsets.cpu().tolist() == np.cumsum([0] + [lens[i] for i in idx]).tolist()
            yield s.bags, s.t, s.e, s.visible

    def direct():
        for idx in groups:
            x = torch.cat([bags[i] for i in idx]).to(torch.bfloat16).cuda()
            lab = torch.tensor([labels[i] for i in idx])
            yield (ops.PackedBags(x, [lens[i] for i in idx]), lab[:, 0].cuda(), lab[:, 1].cuda(),
                   torch.ones(len(idx), dtype=torch.uint8).cuda())

    o1, p1 = run(via_feeder)
    o2, p2 = run(direct)
    for a, b in zip(o1, o2):
        assert torch.equal(a, b)
    for a, b in zip(p1, p2):
        assert torch.equal(a, b)


@pytest.mark.parametrize("case", ["mixed", "all_unlabelled", "hinge"])
def test_semi_supervised_steps_vs_oracle_trainer(case):
    """configs[4]: labelled and unlabelled bags in one step.  Unlabelled bags (label_visible_mask == 0) give the
    discriminator a fake pair only and the generator no reconstruction term (model_handler.py:373-377,465-470); a step
    without any real pair takes the un-batched head path.  fp32 mode against the oracle's restatement of
    _update_disc/_update_gen with torch.optim.Adam, two consecutive steps."""
    from advmil_b200 import ops
    from advmil_b200.step import AdvStep
    Ns = [160, 320, 96, 640, 48]
    B = len(Ns)
    sdG, sdD = O.synth_state_dict(O.G_SHAPES(), 31), O.synth_state_dict(O.D_SHAPES(), 32)
    which = "hinge" if case == "hinge" else "bce"
    tr = O.CpuTrainer(sdG, sdD, which=which)
    G, D = build_G(), build_D()
    G.load_state_dict(sdG)
    D.load_state_dict(sdD)
    eng = AdvStep(G, D, loss_d=which)
    xs = [O.synth_bag(n, 40 + i) for i, n in enumerate(Ns)]
    ts, es = O.synth_labels(B, 9)
    es[1] = 1.0
    vis = [False] * B if case == "all_unlabelled" else [True, True, False, True, False]
    bags = ops.PackedBags.from_list([x.cuda() for x in xs])
    for step in range(2):
        rng = np.random.default_rng(50 + step)
        nd = torch.tensor(rng.uniform(size=(B, 192)), dtype=torch.float32)
        ng = torch.tensor(rng.uniform(size=(B, 192)), dtype=torch.float32)
        mr = [d_masks(n // 16, 128, 300 + 10 * i + step) for i, n in enumerate(Ns)]
        mf = [d_masks(n // 16, 128, 400 + 10 * i + step) for i, n in enumerate(Ns)]
        mg = [g_masks(n, 384, 384, 500 + 10 * i + step) for i, n in enumerate(Ns)]
        ref = tr.step(xs, ts, es, vis, list(nd), list(ng), mr, mf, mg)
        out = eng.step(bags, ts.cuda(), es.cuda(), torch.tensor(vis, dtype=torch.uint8).cuda(), noise_d=nd.cuda(), noise_g=ng.cuda(),
                       masks_d_real=_cat_masks(mr, ["fc1", "ga", "gs", "fc2"]), masks_d_fake=_cat_masks(mf, ["fc1", "ga", "gs", "fc2"]),
                       masks_g=_cat_masks(mg, ["h", "a", "b", "rho", "mlp0"]))
        L = eng.loss_dict(out)
        # hinge with every term active: d(loss)/d(prj_layer.bias) = sum(+1/n_fake) + sum(-1/n_real) = 0 exactly, so its
        # fp32 value is rounding noise whose SIGN Adam turns into a +-lr move of that bias; every score after a D update
        # then carries an offset of up to lr per update (any two correct fp32 evaluations differ there)
        slack = 8e-5 if case == "hinge" else 0.0
        assert_close(out["pred_d"].cpu(), ref["pred_d"].reshape(-1), RTOL, f"pred_d {step}", atol=slack * step * 1e-2)
        assert_close(out["pred_g"].cpu(), ref["pred_g"].reshape(-1), RTOL, f"pred_g {step}", atol=slack * step * 1e-2)
        assert_close(out["f_fake_d"].cpu(), ref["fake_d"].reshape(-1), RTOL, f"fake_d {step}", atol_scale=1e-1, atol=slack * step)
        assert_close(out["f_fake_g"].cpu(), ref["fake_g"].reshape(-1), RTOL, f"fake_g {step}", atol_scale=1e-1, atol=slack * (step + 1))
        tol = 2e-5 + slack * (step + 1)
        assert abs(L["dis_loss"] - ref["dis_loss"]) < tol and abs(L["gen_loss"] - ref["gen_loss"]) < tol
        assert abs(L["t_reg_loss"] - ref["t_reg"]) < tol and abs(L["gen_total_loss"] - ref["total"]) < tol
        if case == "all_unlabelled":
            assert out["f_real"] is None and L["t_reg_loss"] == 0.0
    for k, p in D.named_parameters():
        if not k.endswith(ZERO_GRAD):     # hinge: the per-bag upstream gradients sum to zero, see `slack` above
            assert_close(p.detach().cpu(), tr.sdD[k].detach(), RTOL, "D param " + k, atol=8e-5 * 2 * (1.0 if case == "hinge" else 2e-2))
    for k, p in G.named_parameters():
        if not k.endswith(ZERO_GRAD):
            assert_close(p.detach().cpu(), tr.sdG[k].detach(), RTOL, "G param " + k, atol=8e-5 * 2 * 2e-2)


def test_module_step_with_esat_generator_vs_oracle_trainer():
    """ModuleAdvStep (packed bags, flat buffers, fused Adam with L1/weight decay) with the ESAT generator and the RLIP
    discriminator: two optimiser steps in the fp32 mode against the oracle's restatement of _update_disc/_update_gen over
    the same bags (mixed labelled / unlabelled), injected masks at every dropout site."""
    from advmil_b200 import ops
    from advmil_b200.step import ModuleAdvStep
    C, d = 1024, 384
    Ns = [160, 320, 96, 640]
    B = len(Ns)
    sdG, sdD = O.synth_state_dict(O.G_ESAT_SHAPES(C, d), 131), O.synth_state_dict(O.D_SHAPES(), 132)
    tr = O.CpuTrainer(sdG, sdD, backbone="patch")
    G, D = build_G((C, d, d), mode="patch"), build_D()
    G.load_state_dict(sdG)
    D.load_state_dict(sdD)
    eng = ModuleAdvStep(G, D)
    xs = [O.synth_bag(n, 140 + i, C) for i, n in enumerate(Ns)]
    ts, es = O.synth_labels(B, 133)
    es[0] = 1.0
    vis = [True, True, False, True]
    bags = ops.PackedBags.from_list([x.cuda() for x in xs])
    for step in range(2):
        rng = np.random.default_rng(134 + step)
        nd = torch.tensor(rng.uniform(size=(B, d // 2)), dtype=torch.float32)
        ng = torch.tensor(rng.uniform(size=(B, d // 2)), dtype=torch.float32)
        mr = [d_masks(n // 16, 128, 600 + 10 * i + step) for i, n in enumerate(Ns)]
        mf = [d_masks(n // 16, 128, 700 + 10 * i + step) for i, n in enumerate(Ns)]
        mg = [esat_masks(n // 16, d, 800 + 10 * i + step) for i, n in enumerate(Ns)]
        ref = tr.step(xs, ts, es, vis, list(nd), list(ng), mr, mf, mg)
        mgd = _cat_masks(mg, ["sa", "ff1", "ff2", "ga", "gs", "mlp0"])
        mgd["attn"] = [m["attn"].to(torch.uint8).cuda() for m in mg]
        out = eng.step(bags, ts.cuda(), es.cuda(), torch.tensor(vis, dtype=torch.uint8).cuda(), noise_d=nd.cuda(), noise_g=ng.cuda(),
                       masks_d_real=_cat_masks(mr, ["fc1", "ga", "gs", "fc2"]), masks_d_fake=_cat_masks(mf, ["fc1", "ga", "gs", "fc2"]),
                       masks_g=mgd)
        assert_close(out["pred_d"].cpu(), ref["pred_d"].reshape(-1), RTOL, f"pred_d {step}")
        assert_close(out["pred_g"].cpu(), ref["pred_g"].reshape(-1), RTOL, f"pred_g {step}")
        assert_close(out["f_fake_d"].cpu(), ref["fake_d"].reshape(-1), RTOL, f"fake_d {step}", atol_scale=1e-1)
        assert_close(out["f_fake_g"].cpu(), ref["fake_g"].reshape(-1), RTOL, f"fake_g {step}", atol_scale=1e-1)
        assert abs(float(out["dis_loss"]) - ref["dis_loss"]) < 2e-5 and abs(float(out["gen_loss"]) - ref["gen_loss"]) < 2e-5
        assert abs(float(out["t_reg_loss"]) - ref["t_reg"]) < 2e-5 and abs(float(out["gen_total_loss"]) - ref["total"]) < 2e-5
    for k, p in D.named_parameters():
        if not k.endswith(ZERO_GRAD):
            assert_close(p.detach().cpu(), tr.sdD[k].detach(), RTOL, "D param " + k, atol=8e-5 * 2 * 2e-2)
    for k, p in G.named_parameters():
        if not k.endswith(ZERO_GRAD):
            assert_close(p.detach().cpu(), tr.sdG[k].detach(), RTOL, "G param " + k, atol=8e-5 * 2 * 2e-2)


def test_module_step_with_cluster_generator_vs_oracle_trainer():
    """configs[3]: the full adversarial step with the DeepAttMISL generator (per-cluster mean of relu(phi(x)), gated
    attention over the 8 clusters; one bag with an empty cluster) and the RLIP discriminator through ModuleAdvStep, fp32
    mode, two steps against the oracle trainer."""
    from advmil_b200 import ops
    from advmil_b200.step import ModuleAdvStep
    C, h = 1024, 384
    Ns = [160, 320, 96]
    B = len(Ns)
    sdG, sdD = O.synth_state_dict(O.G_CLUSTER_SHAPES(C, h), 151), O.synth_state_dict(O.D_SHAPES(), 152)
    tr = O.CpuTrainer(sdG, sdD, backbone="cluster")
    G, D = build_G((C, h, h), mode="cluster"), build_D()
    G.load_state_dict(sdG)
    D.load_state_dict(sdD)
    eng = ModuleAdvStep(G, D)
    xs = [O.synth_bag(n, 160 + i, C) for i, n in enumerate(Ns)]
    rng = np.random.default_rng(153)
    cids = [torch.tensor(rng.integers(0, 8, size=n), dtype=torch.float32) for n in Ns]
    cids[1][cids[1] == 5] = 4.0            # cluster 5 of bag 1 is empty (model/backbone.py:114-115)
    ts, es = O.synth_labels(B, 154)
    es[0] = 1.0
    vis = [True, True, True]
    bags = ops.PackedBags.from_list([x.cuda() for x in xs])
    for step in range(2):
        nd = torch.tensor(rng.uniform(size=(B, h // 2)), dtype=torch.float32)
        ng = torch.tensor(rng.uniform(size=(B, h // 2)), dtype=torch.float32)
        mr = [d_masks(n // 16, 128, 900 + 10 * i + step) for i, n in enumerate(Ns)]
        mf = [d_masks(n // 16, 128, 950 + 10 * i + step) for i, n in enumerate(Ns)]
        mg = [{"h": O.synth_masks((8, h), 0.75, 1000 + 10 * i + step), "a": O.synth_masks((8, h), 0.75, 1001 + 10 * i + step),
               "b": O.synth_masks((8, h), 0.75, 1002 + 10 * i + step), "mlp0": O.synth_masks((1, h // 2), 0.4, 1003 + 10 * i + step)}
              for i in range(B)]
        ref = tr.step(xs, ts, es, vis, list(nd), list(ng), mr, mf, mg, exts=cids)
        out = eng.step(bags, ts.cuda(), es.cuda(), torch.tensor(vis, dtype=torch.uint8).cuda(), noise_d=nd.cuda(), noise_g=ng.cuda(),
                       masks_d_real=_cat_masks(mr, ["fc1", "ga", "gs", "fc2"]), masks_d_fake=_cat_masks(mf, ["fc1", "ga", "gs", "fc2"]),
                       masks_g=_cat_masks(mg, ["h", "a", "b", "mlp0"]), ext=torch.cat(cids).cuda())
        assert_close(out["pred_d"].cpu(), ref["pred_d"].reshape(-1), RTOL, f"pred_d {step}")
        assert_close(out["pred_g"].cpu(), ref["pred_g"].reshape(-1), RTOL, f"pred_g {step}")
        assert_close(out["f_fake_g"].cpu(), ref["fake_g"].reshape(-1), RTOL, f"fake_g {step}", atol_scale=1e-1)
        assert abs(float(out["dis_loss"]) - ref["dis_loss"]) < 2e-5 and abs(float(out["gen_total_loss"]) - ref["total"]) < 2e-5
    for k, p in D.named_parameters():
        if not k.endswith(ZERO_GRAD):
            assert_close(p.detach().cpu(), tr.sdD[k].detach(), RTOL, "D param " + k, atol=8e-5 * 2 * 2e-2)
    for k, p in G.named_parameters():
        if not k.endswith(ZERO_GRAD):
            assert_close(p.detach().cpu(), tr.sdG[k].detach(), RTOL, "G param " + k, atol=8e-5 * 2 * 2e-2)


def test_p12_transport_decodes_bit_exactly_on_the_device_and_through_the_feeder():
    """csrc/codec.cu against the host decoder: every bf16 bit pattern class, escapes, and the DeviceFeeder path (the
    device buffer of a p12 step equals the stored bf16 features word for word)."""
    from advmil_b200.dataset.codec import decode_p12_device, encode_bf16_p12
    from advmil_b200.dataset.packed import DeviceFeeder, pack_step
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2048, 1024, generator=g).to(torch.bfloat16)
    x[0, :8] = torch.tensor([0.0, -0.0, float("inf"), float("-inf"), float("nan"), 1e-40, -3e38, 1.0]).to(torch.bfloat16)
    x[7] = (torch.randn(1024, generator=g) * torch.logspace(-30, 30, 1024)).to(torch.bfloat16)
    p = encode_bf16_p12(x)
    out = torch.empty(x.shape, dtype=torch.bfloat16, device="cuda")
    decode_p12_device(p.lo.cuda(), p.hi.cuda(), p.table, p.esc_idx.cuda(), p.esc_exp.cuda(), out)
    assert torch.equal(out.cpu().view(torch.int16), x.view(torch.int16))
    allbits = torch.arange(0, 65536, dtype=torch.int32).to(torch.int16).view(torch.bfloat16).reshape(64, 1024)   # every bf16 word
    pa = encode_bf16_p12(allbits)
    oa = torch.empty(allbits.shape, dtype=torch.bfloat16, device="cuda")
    decode_p12_device(pa.lo.cuda(), pa.hi.cuda(), pa.table, pa.esc_idx.cuda(), pa.esc_exp.cuda(), oa)
    assert torch.equal(oa.cpu().view(torch.int16), allbits.view(torch.int16))
    lens = [640, 16, 1040, 352]
    bags = [torch.randn(n, 1024, generator=g) for n in lens]
    steps = [pack_step(bags, [(0.5, 1.0)] * 4, dtype=torch.bfloat16).pack12() for _ in range(3)]
    assert steps[0].nbytes < 0.76 * steps[0].x.numel() * 2
    for s in DeviceFeeder(steps, device="cuda"):
        assert torch.equal(s.bags.x.cpu().view(torch.int16), steps[0].x.view(torch.int16))
        assert s.bags.offsets.cpu().tolist() == [0, 640, 656, 1696, 2048]


@pytest.mark.parametrize("transport", ["p12", "vl"])
def test_p12_packed_file_through_the_feeder(tmp_path, transport):
    """A split stored in a transport form (dataset/packed_file.py, transport="p12" / "vl"): the feeder copies the planes
    (and Huffman streams) as stored and the device sees exactly the bf16 features of the raw bf16 file, step after step."""
    from advmil_b200.dataset.packed import DeviceFeeder, group_steps
    from advmil_b200.dataset.packed_file import PackedFile, write_packed
    g = torch.Generator().manual_seed(13)
    lens = [64, 320, 16, 160, 96, 48, 640, 32]
    bags = [torch.randn(n, 1024, generator=g) for n in lens]
    labels = [(0.1 + 0.1 * i, float(i % 2 == 0)) for i in range(len(lens))]
    write_packed(str(tmp_path / "p12.advmil"), iter(bags), labels, dtype=torch.bfloat16, transport=transport)
    pf = PackedFile(str(tmp_path / "p12.advmil"))
    groups = group_steps(len(pf), 4)
    for s, idx in zip(DeviceFeeder((pf.step(i) for i in groups), device="cuda"), groups):
        want = torch.cat([bags[i] for i in idx]).to(torch.bfloat16)
        assert torch.equal(s.bags.x.cpu().view(torch.int16), want.view(torch.int16))
        assert s.t.cpu().tolist() == pytest.approx([labels[i][0] for i in idx])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cindex_and_sampled_distributions_over_447_patients(precision):
    """configs[0]'s evaluation: MyHandler.test_model over 447 patients (NLST size), 1 + 30 generator samples per bag, lower
    median, Harrell's C (model/model_handler.py:598-643, eval/cindex.py).  The CUDA path replays the oracle's CPU noise
    stream; fp32 mode: every sampled time within 1e-5 and the same C-index; bf16 mode: sampled times within 2e-2 and the
    C-index within 0.005 (north_star)."""
    from advmil_b200 import ops
    from advmil_b200.eval.cindex import concordance_index
    from advmil_b200.step import sample_inference
    P, S = 447, 30
    rng = np.random.default_rng(21)
    Ns = [16 * int(rng.integers(2, 12)) for _ in range(P)]
    sdG, sdD = O.synth_state_dict(O.G_SHAPES(), 23), O.synth_state_dict(O.D_SHAPES(), 24)
    G, D = build_G().eval(), build_D().eval()
    G.load_state_dict(sdG)
    D.load_state_dict(sdD)
    xs = [O.synth_bag(n, 3000 + i) for i, n in enumerate(Ns)]
    ts, es = O.synth_labels(P, 25)
    es[0] = 1.0
    torch.manual_seed(99)
    res = sample_inference(G, D, ops.PackedBags.from_list([x.cuda() for x in xs]), times_test_sample=S, precision=precision)
    torch.manual_seed(99)
    first = torch.rand(P, 192)
    draws = [torch.rand(P, 192) for _ in range(S)]
    want_dist = torch.stack([O.sample_times(sdG, x, [d_[b:b + 1] for d_ in draws]) for b, x in enumerate(xs)])    # [P, S]
    want_avg = O.lower_median(want_dist, dim=1)
    tol = 1e-5 if precision == "fp32" else 2e-2
    assert_close(res["dist_y_hat"][:, :, 0].cpu(), want_dist, tol, "sampled times")
    y_true = np.stack([ts.numpy(), es.numpy()], axis=1)
    ci_ref = O.concordance_index(ts.numpy(), es.numpy(), want_avg.numpy())
    ci = concordance_index(y_true, res["avg_y_hat"].cpu().numpy().reshape(-1, 1))
    if precision == "fp32":
        assert abs(ci - ci_ref) <= 1e-4, (ci, ci_ref)      # a median may pick the neighbouring sample where two are 1e-7 apart
    else:
        assert abs(ci - ci_ref) <= 0.005, (ci, ci_ref)
    assert 0.0 <= ci <= 1.0 and first.shape == (P, 192)


def test_fused_step_many_small_bags_vs_oracle_trainer():
    """64 bags of 16..320 rows in one fused step (bags smaller than one GEMM tile, several bags per row chunk, four times
    the benchmark's 16 bags, labelled and unlabelled mixed): fp32 mode, one D+G step with injected masks."""
    from advmil_b200 import ops
    from advmil_b200.step import AdvStep
    rng = np.random.default_rng(171)
    Ns = [16 * int(rng.integers(1, 21)) for _ in range(64)]
    B = len(Ns)
    sdG, sdD = O.synth_state_dict(O.G_SHAPES(), 172), O.synth_state_dict(O.D_SHAPES(), 173)
    tr = O.CpuTrainer(sdG, sdD)
    G, D = build_G(), build_D()
    G.load_state_dict(sdG)
    D.load_state_dict(sdD)
    eng = AdvStep(G, D)
    xs = [O.synth_bag(n, 5000 + i) for i, n in enumerate(Ns)]
    ts, es = O.synth_labels(B, 174)
    es[0] = 1.0
    vis = [bool(v) for v in rng.uniform(size=B) < 0.6]
    vis[0] = True
    nd = torch.tensor(rng.uniform(size=(B, 192)), dtype=torch.float32)
    ng = torch.tensor(rng.uniform(size=(B, 192)), dtype=torch.float32)
    mr = [d_masks(n // 16, 128, 7000 + 10 * i) for i, n in enumerate(Ns)]
    mf = [d_masks(n // 16, 128, 8000 + 10 * i) for i, n in enumerate(Ns)]
    mg = [g_masks(n, 384, 384, 9000 + 10 * i) for i, n in enumerate(Ns)]
    ref = tr.step(xs, ts, es, vis, list(nd), list(ng), mr, mf, mg)
    out = eng.step(ops.PackedBags.from_list([x.cuda() for x in xs]), ts.cuda(), es.cuda(), torch.tensor(vis, dtype=torch.uint8).cuda(),
                   noise_d=nd.cuda(), noise_g=ng.cuda(), masks_d_real=_cat_masks(mr, ["fc1", "ga", "gs", "fc2"]),
                   masks_d_fake=_cat_masks(mf, ["fc1", "ga", "gs", "fc2"]), masks_g=_cat_masks(mg, ["h", "a", "b", "rho", "mlp0"]))
    L = eng.loss_dict(out)
    assert_close(out["pred_d"].cpu(), ref["pred_d"].reshape(-1), RTOL, "pred_d")
    assert_close(out["pred_g"].cpu(), ref["pred_g"].reshape(-1), RTOL, "pred_g")
    assert_close(out["f_fake_d"].cpu(), ref["fake_d"].reshape(-1), RTOL, "fake_d", atol_scale=1e-1)
    assert_close(out["f_fake_g"].cpu(), ref["fake_g"].reshape(-1), RTOL, "fake_g", atol_scale=1e-1)
    assert abs(L["dis_loss"] - ref["dis_loss"]) < 2e-5 and abs(L["gen_loss"] - ref["gen_loss"]) < 2e-5
    assert abs(L["t_reg_loss"] - ref["t_reg"]) < 2e-5 and abs(L["gen_total_loss"] - ref["total"]) < 2e-5
    for k, p in D.named_parameters():
        if not k.endswith(ZERO_GRAD):
            assert_close(p.detach().cpu(), tr.sdD[k].detach(), RTOL, "D param " + k, atol=8e-5 * 2e-2)
    for k, p in G.named_parameters():
        if not k.endswith(ZERO_GRAD):
            assert_close(p.detach().cpu(), tr.sdG[k].detach(), RTOL, "G param " + k, atol=8e-5 * 2e-2)


def test_vl_transport_decodes_bit_exactly_on_the_device_and_through_the_feeder():
    """The entropy-coded transport form: the device decoder equals the stored bf16 words for Gaussian features, for every bit
    pattern and for zeros-heavy features, and a packvl() step arrives through the DeviceFeeder word for word."""
    from advmil_b200.dataset.codec import decode_vl_device, encode_bf16_vl
    from advmil_b200.dataset.packed import DeviceFeeder, pack_step
    torch.manual_seed(5)
    cases = [torch.randn(256 * 4096).to(torch.bfloat16),
             torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.bfloat16).repeat(2),
             torch.relu(torch.randn(32 * 4096)).mul(0.5).to(torch.bfloat16), torch.zeros(4096, dtype=torch.bfloat16)]
    for x in cases:
        p = encode_bf16_vl(x)
        out = torch.empty(x.numel(), dtype=torch.bfloat16, device="cuda")
        decode_vl_device(p.lo.cuda(), p.stream.cuda(), p.sbase.cuda(), p.loff.cuda(), p, p.esc_idx.cuda(), p.esc_exp.cuda(), out)
        assert torch.equal(out.cpu().view(torch.int16), x.view(torch.int16))
    xs = [torch.randn(n, 1024) for n in (320, 1600, 128)]
    labels = [(0.3, 1.0), (0.5, 0.0), (0.9, 1.0)]
    steps = [pack_step(xs, labels, dtype=torch.bfloat16).packvl() for _ in range(3)]
    assert steps[0].nbytes < 0.70 * steps[0].x.numel() * 2
    for st, dv in zip(steps, DeviceFeeder(steps, device="cuda", depth=2)):
        torch.cuda.current_stream().synchronize()
        assert torch.equal(dv.bags.x.cpu().view(torch.int16), st.x.view(torch.int16))


def _esat_nets_without_dropout(C, d, seed_g, seed_d):
    sdG, sdD = O.synth_state_dict(O.G_ESAT_SHAPES(C, d), seed_g), O.synth_state_dict(O.D_SHAPES(), seed_d)
    G, D = build_G((C, d, d), mode="patch"), build_D()
    G.load_state_dict(sdG)
    D.load_state_dict(sdD)
    G.backbone.p = G.backbone.pool.p = 0.0          # the C step draws its own dropout bits: compare the deterministic arithmetic
    G.p_head = 0.0
    D.net_pair_one.p = 0.0
    return sdG, sdD, G, D


def test_c_fused_esat_step_vs_oracle_trainer_and_module_step():
    """EsatAdvStep (advmil_adv_step_esat_disc / _gen: the ESAT generator's adversarial step issued from C) in the fp32 mode:
    two optimiser steps against the oracle's restatement of _update_disc/_update_gen (dropout probabilities zeroed) and
    against ModuleAdvStep on the same inputs; mixed labelled / unlabelled bags."""
    from advmil_b200 import ops
    from advmil_b200.step import EsatAdvStep, ModuleAdvStep
    C, d = 1024, 384
    Ns = [160, 320, 96, 640]
    B = len(Ns)
    sdG, sdD, G, D = _esat_nets_without_dropout(C, d, 171, 172)
    _, _, G2, D2 = _esat_nets_without_dropout(C, d, 171, 172)
    tr = O.CpuTrainer(sdG, sdD, backbone="patch")
    eng, mod = EsatAdvStep(G, D), ModuleAdvStep(G2, D2)
    xs = [O.synth_bag(n, 180 + i, C) for i, n in enumerate(Ns)]
    ts, es = O.synth_labels(B, 173)
    es[0] = 1.0
    vis = [True, True, False, True]
    bags = ops.PackedBags.from_list([x.cuda() for x in xs])
    for step in range(2):
        rng = np.random.default_rng(174 + step)
        nd = torch.tensor(rng.uniform(size=(B, d // 2)), dtype=torch.float32)
        ng = torch.tensor(rng.uniform(size=(B, d // 2)), dtype=torch.float32)
        ref = tr.step(xs, ts, es, vis, list(nd), list(ng))
        args = (bags, ts.cuda(), es.cuda(), torch.tensor(vis, dtype=torch.uint8).cuda())
        out = eng.step(*args, noise_d=nd.cuda(), noise_g=ng.cuda())
        om = mod.step(*args, noise_d=nd.cuda(), noise_g=ng.cuda())
        L = eng.loss_dict(out)
        assert_close(out["pred_d"].cpu(), ref["pred_d"].reshape(-1), RTOL, f"pred_d {step}")
        assert_close(out["pred_g"].cpu(), ref["pred_g"].reshape(-1), RTOL, f"pred_g {step}")
        assert_close(out["f_fake_d"].cpu(), ref["fake_d"].reshape(-1), RTOL, f"fake_d {step}", atol_scale=1e-1)
        assert_close(out["f_fake_g"].cpu(), ref["fake_g"].reshape(-1), RTOL, f"fake_g {step}", atol_scale=1e-1)
        assert abs(L["dis_loss"] - ref["dis_loss"]) < 2e-5 and abs(L["gen_loss"] - ref["gen_loss"]) < 2e-5
        assert abs(L["t_reg_loss"] - ref["t_reg"]) < 2e-5 and abs(L["gen_total_loss"] - ref["total"]) < 2e-5
        assert_close(out["pred_g"].cpu(), om["pred_g"].cpu(), RTOL, f"pred_g vs module step {step}")
    for k, p in D.named_parameters():
        if not k.endswith(ZERO_GRAD):
            assert_close(p.detach().cpu(), tr.sdD[k].detach(), RTOL, "D param " + k, atol=8e-5 * 2 * 2e-2)
    for (k, p), p2 in zip(G.named_parameters(), G2.parameters()):
        if not k.endswith(ZERO_GRAD):
            assert_close(p.detach().cpu(), tr.sdG[k].detach(), RTOL, "G param " + k, atol=8e-5 * 2 * 2e-2)
            assert_close(p.detach().cpu(), p2.detach().cpu(), RTOL, "G param vs module step " + k, atol=8e-5 * 2 * 2e-2)


def test_c_fused_esat_step_bf16_with_dropout_runs_and_learns():
    """bf16 mode with the in-kernel dropout of every site: finite losses, parameters move, and the reconstruction loss
    falls over a few steps on a fixed batch (the full path: tcgen05 attention forward / backward, chain kernels, Adam)."""
    from advmil_b200 import ops
    from advmil_b200.step import EsatAdvStep
    C, d = 1024, 384
    Ns = [320, 640, 160, 2048]
    sdG, sdD = O.synth_state_dict(O.G_ESAT_SHAPES(C, d), 191), O.synth_state_dict(O.D_SHAPES(), 192)
    G, D = build_G((C, d, d), mode="patch"), build_D()
    G.load_state_dict(sdG)
    D.load_state_dict(sdD)
    torch.manual_seed(1234)              # the dropout seeds and the generator noise come from torch's default CPU generator
    eng = EsatAdvStep(G, D, precision="bf16", lr_g=1e-3)
    bags = ops.PackedBags.from_list([O.synth_bag(n, 200 + i, C).cuda() for i, n in enumerate(Ns)])
    t = torch.tensor([0.2, 0.5, 0.7, 0.9], device="cuda")
    e = torch.ones(4, device="cuda")
    vis = torch.ones(4, dtype=torch.uint8, device="cuda")
    nz = torch.rand(4, d // 2, device="cuda")
    p0 = eng.G.flat.clone()
    hist = []
    for _ in range(30):
        L = eng.loss_dict(eng.step(bags, t, e, vis, noise_d=nz, noise_g=nz))
        assert all(np.isfinite(v) for v in L.values())
        hist.append(L["t_reg_loss"])
    assert float((eng.G.flat - p0).abs().max()) > 0
    assert float(np.mean(hist[-5:])) < float(np.mean(hist[:3])), hist
