"""Parity at BASELINE.json's full sizes (configs[1]: 16 bags of 16384 x 1024 per step; configs[2]: ragged bags up to
100k rows): the CUDA step against the oracle on the same seeded inputs (a 16k-row bag takes the oracle ~0.1 s), plus the
size-independent properties the path offers — bag independence (packed == alone), additivity of gradients over bag
shards under global-count normalisation (the data-parallel contract), permutation invariance of MIL pooling over
instances and of RLIP over regions, run-to-run determinism, and the closed form of the region index map."""
import numpy as np
import pytest
import torch

from oracle import advmil_oracle as O
from tests.util import assert_close, build_D, build_G, d_masks, g_masks

pytestmark = pytest.mark.gpu
ZERO_GRAD = ("pool.fc2.bias", "attention_c.bias")
N_FULL, B_FULL = 16384, 16


def _cat(per_bag, keys):
    return {k: torch.cat([m[k] for m in per_bag], dim=0).to(torch.uint8).contiguous().cuda() for k in keys}


def _engine(precision, sdG, sdD, **kw):
    from advmil_b200.step import AdvStep
    G, D = build_G(), build_D()
    G.load_state_dict(sdG)
    D.load_state_dict(sdD)
    return AdvStep(G, D, precision=precision, **kw), G, D


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "bf16"])
def test_full_size_step_vs_oracle(precision):
    """One D step + G step + both Adam updates on 16 x 16384 x 1024 with injected dropout masks against the oracle's
    restatement of _update_disc/_update_gen (model/model_handler.py:349-498).  fp32 mode: rtol 1e-5 (SURVEY.md tier
    framing); bf16 mode: 2e-2 on outputs and losses."""
    from advmil_b200 import ops
    Ns = [N_FULL] * B_FULL
    sdG, sdD = O.synth_state_dict(O.G_SHAPES(), 61), O.synth_state_dict(O.D_SHAPES(), 62)
    xs = [O.synth_bag(n, 600 + i) for i, n in enumerate(Ns)]
    ts, es = O.synth_labels(B_FULL, 63)
    es[0] = 1.0
    vis = [True] * B_FULL
    rng = np.random.default_rng(64)
    nd = torch.tensor(rng.uniform(size=(B_FULL, 192)), dtype=torch.float32)
    ng = torch.tensor(rng.uniform(size=(B_FULL, 192)), dtype=torch.float32)
    mr = [d_masks(n // 16, 128, 700 + 10 * i) for i, n in enumerate(Ns)]
    mf = [d_masks(n // 16, 128, 800 + 10 * i) for i, n in enumerate(Ns)]
    mg = [g_masks(n, 384, 384, 900 + 10 * i) for i, n in enumerate(Ns)]
    ref = O.CpuTrainer(sdG, sdD).step(xs, ts, es, vis, list(nd), list(ng), mr, mf, mg)
    eng, G, D = _engine(precision, sdG, sdD)
    bags = ops.PackedBags.from_list([x.cuda() for x in xs])
    out = eng.step(bags, ts.cuda(), es.cuda(), torch.tensor(vis, dtype=torch.uint8).cuda(), noise_d=nd.cuda(), noise_g=ng.cuda(),
                   masks_d_real=_cat(mr, ["fc1", "ga", "gs", "fc2"]), masks_d_fake=_cat(mf, ["fc1", "ga", "gs", "fc2"]),
                   masks_g=_cat(mg, ["h", "a", "b", "rho", "mlp0"]))
    L = eng.loss_dict(out)
    exact = precision in ("fp32", "tf32x3")          # tf32x3: the exact-parity mode on the tensor cores (split tf32)
    tol = 1e-5 if exact else 2e-2
    assert_close(out["pred_d"].cpu(), ref["pred_d"].reshape(-1), tol, "pred_d")
    assert_close(out["pred_g"].cpu(), ref["pred_g"].reshape(-1), tol, "pred_g")
    assert_close(out["f_fake_d"].cpu(), ref["fake_d"].reshape(-1), tol, "fake_d", atol_scale=1e-1)
    assert_close(out["f_fake_g"].cpu(), ref["fake_g"].reshape(-1), tol, "fake_g", atol_scale=1e-1)
    ltol = 2e-5 if exact else 2e-2
    assert abs(L["dis_loss"] - ref["dis_loss"]) < ltol and abs(L["gen_loss"] - ref["gen_loss"]) < ltol
    assert abs(L["t_reg_loss"] - ref["t_reg"]) < ltol and abs(L["gen_total_loss"] - ref["total"]) < ltol
    if not exact:
        return                     # bf16: gradients and post-Adam parameters in tests/test_gpu_bf16_step.py
    # Gradients (fp32 / split-tf32 modes).  Every weight gradient here is a sum over 262,144 rows in which real and fake pair terms of
    # opposite sign cancel: two correct fp32 evaluations differ by the rounding of the cancelled partial sums, not by 1e-5
    # of the result.  The yardstick is therefore the oracle re-run in float64: the CUDA path must be as close to it as
    # the reference's own fp32 arithmetic is (factor 4), or within rtol 1e-5.
    f64 = lambda sd: {k: v.double() for k, v in sd.items()}                                      # noqa: E731
    ref64 = O.CpuTrainer(f64(sdG), f64(sdD)).step([x.double() for x in xs], ts.double(), es.double(), vis, list(nd.double()),
                                                  list(ng.double()), mr, mf, mg)

    def check(name, got, r32, r64):
        got, r32, r64 = got.detach().double().cpu().reshape(-1), r32.double().reshape(-1), r64.reshape(-1)
        scale = float(r64.abs().max())
        ours, theirs = float((got - r64).abs().max()), float((r32 - r64).abs().max())
        # split tf32: products are fp32-grade but the tensor core's accumulation truncates (stage tests: 4e-6 of the largest
        # entry against 1-2e-6 for an fp32 FFMA chain), hence the wider factor
        factor = 4.0 if precision == "fp32" else 10.0
        assert ours <= max(factor * theirs, 1e-5 * scale), f"{name}: |cuda - f64| {ours:.3e} vs |ref fp32 - f64| {theirs:.3e} (scale {scale:.3e})"

    pos = {id(t): i for i, t in enumerate(eng.dparams) if t is not None}
    for k, p in D.named_parameters():
        if not k.endswith(ZERO_GRAD):
            check("D grad " + k, eng.dgrads[pos[id(p)]], ref["d_grads"][k], ref64["d_grads"][k])
    pos = {id(t): i for i, t in enumerate(eng.gparams) if t is not None}
    for k, p in G.named_parameters():
        if not k.endswith(ZERO_GRAD):
            full = eng.ggrads[pos[id(p)]].cpu() + 1e-5 * torch.sign(sdG[k])     # loss_reg_l1 is folded into Adam
            check("G grad " + k, full, ref["g_grads"][k], ref64["g_grads"][k])


def _run_step(precision, sdG, sdD, xs, ts, es, vis, nd, ng, seeds, counts=None):
    """One in-kernel-dropout step with pinned dropout seeds; returns outputs and the flat gradient buffers."""
    from advmil_b200 import ops
    import advmil_b200.step as step_mod
    it = iter(seeds)
    saved, step_mod.next_dropout_seed = step_mod.next_dropout_seed, (lambda: next(it))
    try:
        eng, G, D = _engine(precision, sdG, sdD)
        out = eng.step(ops.PackedBags.from_list(xs), ts, es, vis, noise_d=nd, noise_g=ng, global_counts=counts)
        torch.cuda.synchronize()
        return out, eng.D.grad.clone(), eng.G.grad.clone(), eng.G.flat.clone(), eng.D.flat.clone()
    finally:
        step_mod.next_dropout_seed = saved


def _inputs(Ns, seed, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    xs = [torch.randn(n, 1024, device="cuda", generator=g).to(dtype) for n in Ns]
    ts, es = O.synth_labels(len(Ns), seed)
    es[0] = 1.0
    rng = np.random.default_rng(seed)
    nd = torch.tensor(rng.uniform(size=(len(Ns), 192)), dtype=torch.float32).cuda()
    ng = torch.tensor(rng.uniform(size=(len(Ns), 192)), dtype=torch.float32).cuda()
    return xs, ts.cuda(), es.cuda(), torch.ones(len(Ns), dtype=torch.uint8).cuda(), nd, ng


def test_full_size_step_is_deterministic():
    """Same inputs, same seeds: every output, gradient and updated parameter is bitwise identical (no atomics in any
    reduction; the side-stream overlap does not change results)."""
    sdG, sdD = O.synth_state_dict(O.G_SHAPES(), 71), O.synth_state_dict(O.D_SHAPES(), 72)
    xs, ts, es, vis, nd, ng = _inputs([N_FULL] * B_FULL, 73, torch.bfloat16)
    a = _run_step("bf16", sdG, sdD, xs, ts, es, vis, nd, ng, [5, 6])
    b = _run_step("bf16", sdG, sdD, xs, ts, es, vis, nd, ng, [5, 6])
    for k in ("pred_d", "pred_g", "f_d", "f_fake_g", "losses"):
        assert torch.equal(a[0][k], b[0][k]), k
    for u, v in zip(a[1:], b[1:]):
        assert torch.equal(u, v)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@torch.no_grad()
def test_ragged_bags_packed_equals_alone(precision):
    """configs[2]: bags of 16 .. 100,000 rows packed into one step.  Bags are independent units: the generator's and the
    discriminator's per-bag outputs in the packed eval pass equal those of each bag run alone (same kernels, different
    tile/segment boundaries -> float reassociation only)."""
    from advmil_b200 import ops
    Ns = [100000, 16, 1024, 50000 - 50000 % 16, 4096, 16384, 48, 30000 - 30000 % 16]
    sdG, sdD = O.synth_state_dict(O.G_SHAPES(), 81), O.synth_state_dict(O.D_SHAPES(), 82)
    G, D = build_G().eval(), build_D().eval()
    G.load_state_dict(sdG)
    D.load_state_dict(sdD)
    xs, ts, _, _, nd, _ = _inputs(Ns, 83, torch.bfloat16 if precision == "bf16" else torch.float32)
    P = ops.PRECISIONS[precision]
    packed = ops.PackedBags.from_list(xs)
    acts = ops.generator_forward(G.config(), G.gen_params(), packed, None, nd, train=False, precision=P, save=False)
    pred = acts["pred"].reshape(-1)
    import advmil_b200
    advmil_b200.set_precision(precision)
    try:
        f = D.forward_packed(packed, ts.reshape(-1, 1)).reshape(-1)
        tol = 1e-5 if precision == "fp32" else 2e-3
        for i, x in enumerate(xs):
            one = ops.PackedBags.from_list([x])
            p1 = ops.generator_forward(G.config(), G.gen_params(), one, None, nd[i:i + 1], train=False, precision=P, save=False)["pred"]
            f1 = D.forward_packed(one, ts[i:i + 1].reshape(1, 1))
            assert_close(pred[i:i + 1].cpu(), p1.reshape(-1).cpu(), tol, f"pred bag {i} ({Ns[i]} rows)")
            assert_close(f[i:i + 1].cpu(), f1.reshape(-1).cpu(), tol, f"D out bag {i} ({Ns[i]} rows)", atol_scale=1e-1)
    finally:
        advmil_b200.set_precision("fp32")


def test_full_size_gradients_are_additive_over_bag_shards():
    """The data-parallel contract at full size on one device: with the loss terms normalised by GLOBAL counts, the
    gradients of the 16-bag step equal the sum of the gradients of its two 8-bag shards (what the NCCL all-reduce adds
    up), for the D step and the G step.  lr_d = 0 keeps D fixed between the phases so the G-step gradients of the three
    runs see the same discriminator."""
    sdG, sdD = O.synth_state_dict(O.G_SHAPES(), 91), O.synth_state_dict(O.D_SHAPES(), 92)
    Ns = [N_FULL] * B_FULL
    xs, ts, es, vis, nd, ng = _inputs(Ns, 93)
    n_real = float(((es == 1) & (vis != 0)).sum())
    counts = (n_real, float(B_FULL), float(B_FULL))
    # in-kernel dropout bits are keyed by a row's position in the packed step, which differs between the whole and a shard:
    # all three runs therefore inject all-ones keep masks (train-mode scaling stays on)
    ones_g = lambda n: {"h": torch.ones(n, 384), "a": torch.ones(n, 384), "b": torch.ones(n, 384),   # noqa: E731
                        "rho": torch.ones(1, 384), "mlp0": torch.ones(1, 192)}
    ones_d = lambda r: {"fc1": torch.ones(r, 64), "ga": torch.ones(r, 128), "gs": torch.ones(r, 128), "fc2": torch.ones(1, 64)}  # noqa: E731

    def run(idx):
        from advmil_b200 import ops
        eng, G, D = _engine("fp32", sdG, sdD, lr_d=0.0)
        sel = torch.tensor(idx).cuda()
        md = _cat([ones_d(Ns[i] // 16) for i in idx], ["fc1", "ga", "gs", "fc2"])
        mgm = _cat([ones_g(Ns[i]) for i in idx], ["h", "a", "b", "rho", "mlp0"])
        eng.step(ops.PackedBags.from_list([xs[i] for i in idx]), ts[sel], es[sel], vis[sel], noise_d=nd[sel], noise_g=ng[sel],
                 masks_d_real=md, masks_d_fake=md, masks_g=mgm, global_counts=counts)
        torch.cuda.synchronize()
        return eng.D.grad.clone(), eng.G.grad.clone()

    dW, gW = run(list(range(B_FULL)))
    dA, gA = run(list(range(0, B_FULL // 2)))
    dB, gB = run(list(range(B_FULL // 2, B_FULL)))
    # additive up to fp32 reassociation of the sums over bags (real and fake pair terms of opposite sign cancel in D)
    assert_close((dA + dB).cpu(), dW.cpu(), 1e-4, "D grads", atol=2.0 ** -23 * 16 * max(float(dA.abs().max()), float(dW.abs().max())))
    assert_close((gA + gB).cpu(), gW.cpu(), 1e-4, "G grads", atol=2.0 ** -23 * 16 * float(gW.abs().max()))


@torch.no_grad()
def test_instance_and_region_permutation_invariance_full_size():
    """MIL pooling is a set function: permuting the instances of a bag leaves G's prediction unchanged; permuting whole
    16-row regions (and rows inside a region) leaves D's RLIP output unchanged (mean over regions of <fi_r, ht> plus
    attention pooling over regions).  One 16384-row bag, fp32 mode, float reassociation only."""
    from advmil_b200 import ops
    sdG, sdD = O.synth_state_dict(O.G_SHAPES(), 101), O.synth_state_dict(O.D_SHAPES(), 102)
    G, D = build_G().eval(), build_D().eval()
    G.load_state_dict(sdG)
    D.load_state_dict(sdD)
    g = torch.Generator(device="cuda").manual_seed(103)
    x = torch.randn(N_FULL, 1024, device="cuda", generator=g)
    noise = torch.rand(1, 192, device="cuda", generator=g)
    t = torch.tensor([[0.37]], device="cuda")
    perm_rows = torch.randperm(N_FULL, device="cuda", generator=g)
    R = N_FULL // 16
    perm_reg = torch.randperm(R, device="cuda", generator=g)
    inner = torch.stack([torch.randperm(16, device="cuda", generator=g) for _ in range(8)])[torch.arange(R, device="cuda") % 8]
    rows_by_region = (perm_reg[:, None] * 16 + inner).reshape(-1)

    def gen(xx):
        return ops.generator_forward(G.config(), G.gen_params(), ops.PackedBags.from_list([xx]), None, noise, train=False,
                                     precision=ops.FP32, save=False)["pred"].reshape(-1).cpu()

    assert_close(gen(x[perm_rows].contiguous()), gen(x), 1e-5, "G under instance permutation")
    f0 = D.forward_packed(ops.PackedBags.from_list([x]), t).reshape(-1).cpu()
    f1 = D.forward_packed(ops.PackedBags.from_list([x[rows_by_region].contiguous()]), t).reshape(-1).cpu()
    assert_close(f1, f0, 1e-5, "D under region permutation", atol_scale=1e-1)


def test_region_index_map_closed_form_at_maximum_size():
    """tools/big_to_small_patching.py:40-46,59-76 at 100,000 level-2 patches (1.6 M level-1 rows): row 16k + 4j + i is
    child (i, j) of parent k — integer exact, checked against the closed form and a checksum; region_of_rows == n // 16
    over the largest packed step (16 x 100k rows)."""
    from advmil_b200 import ops
    M = 100000
    rng = np.random.default_rng(5)
    coords = torch.tensor(rng.integers(0, 200000, size=(M, 2)), dtype=torch.int32).cuda()
    out = ops.region_index_map(coords, 256)
    want = O.region_index_map(coords.cpu().numpy(), 256)
    got = out.cpu().numpy()
    assert got.shape == (16 * M, 2) and got.dtype == want.dtype
    assert np.array_equal(got, want)
    k = np.arange(16 * M)
    closed = coords.cpu().numpy()[k // 16].astype(np.float64) + np.stack([(k % 4) * 256, ((k % 16) // 4) * 256], axis=1)
    assert np.array_equal(got.astype(np.float64), closed)
    rows = 16 * 100000
    reg = ops.region_of_rows(rows, device="cuda").cpu().numpy()          # [rows, 3] = (region, j, i)
    n = np.arange(rows)
    assert np.array_equal(reg[:, 0], n // 16) and np.array_equal(reg[:, 1], (n % 16) // 4) and np.array_equal(reg[:, 2], n % 4)
    assert int(reg[:, 0].astype(np.int64).sum()) == int((n // 16).sum())


@pytest.mark.parametrize("form", ["vl", "p12"])
def test_transport_forms_are_bit_exact_at_the_full_step_size(form):
    """BASELINE.json's step size (16 x 16384 x 1024 bf16 = 268 M elements) through the packed loader's transport forms:
    encode on the host, copy + decode on the device (DeviceFeeder), compare every word with the stored bf16 features -- for
    Gaussian features with injected zeros, denormals, infinities and NaN payloads (escapes)."""
    from advmil_b200.dataset.packed import DeviceFeeder, PinnedStep
    torch.manual_seed(11)
    rows, C = 16 * 16384, 1024
    x = torch.randn(rows, C).to(torch.bfloat16)
    v = x.view(torch.int16).reshape(-1)
    idx = torch.randint(0, v.numel(), (4096,))
    v[idx] = torch.randint(-32768, 32767, (4096,), dtype=torch.int32).to(torch.int16)        # arbitrary bit patterns: escapes
    v[:64] = 0
    st = PinnedStep(x=x, lengths=[16384] * 16, t=torch.rand(16), e=torch.ones(16), idx=torch.arange(16, dtype=torch.int32),
                    visible=torch.ones(16, dtype=torch.uint8))
    st = st.packvl() if form == "vl" else st.pack12()
    bits = 8.0 * st.nbytes / x.numel()
    assert (bits < 11.1) if form == "vl" else (bits < 12.1)
    for dv in DeviceFeeder([st], device="cuda", depth=2):
        torch.cuda.synchronize()
        assert torch.equal(dv.bags.x.view(torch.int16).cpu(), x.view(torch.int16))
