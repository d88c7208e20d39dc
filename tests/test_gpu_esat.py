"""GPU parity of the ESAT generator (bcb_mode 'patch': DualTrans_HS, reference model/backbone.py:171-196) through the
drop-in modules: fixtures written by the live reference (tests/golden/g_esat_*.npz), packed ragged bags against the
oracle, and the in-kernel dropout generator against the oracle fed with the same bits."""
import numpy as np
import pytest
import torch

import advmil_b200
from advmil_b200 import ops
from oracle import advmil_oracle as O
from tests.util import assert_close, build_G, condition_esat_case, esat_masks, golden, sub

pytestmark = pytest.mark.gpu
ZERO_GRAD = ("pool.fc2.bias",)      # mathematically zero (softmax shift invariance)


def _dev_masks(m):
    out = {k: v.to(torch.uint8).contiguous().cuda() for k, v in m.items() if k != "attn"}
    out["attn"] = [m["attn"].to(torch.uint8).cuda()]
    return out


@pytest.mark.parametrize("name", ["g_esat_eval_full", "g_esat_train_full", "g_esat_train_small", "g_esat_eval_small"])
def test_esat_generator_vs_reference_golden(name):
    """Generator(x, coord) with the ESAT backbone, single bag through the reference's forward signature, fp32 mode:
    outputs 1e-5, gradients 1e-5 (+ one fp32 ulp of the largest gradient for cancelling sums)."""
    g = golden(name)
    C, d, N, train, seed, with_coord = [int(v) for v in g["cfg"]]
    sd = O.synth_state_dict(O.G_ESAT_SHAPES(C, d), seed)
    G = build_G((C, d, d), mode="patch")
    G.load_state_dict(sd)
    G.train(bool(train))
    x = O.synth_bag(N, seed, C)
    noise = torch.tensor(np.random.default_rng(seed + 7).uniform(size=(1, d // 2)), dtype=torch.float32)
    G.draw_noise = lambda nb, dev, zero: [None, noise.to(dev)]
    coord = torch.tensor(g["coord"]).cuda().unsqueeze(0) if with_coord else None
    if train:
        G._inject_masks = _dev_masks(esat_masks(N // 16, d, seed * 10))
    inter = {}
    pred = G(x.cuda().unsqueeze(0), coord)
    pred.sum().backward()
    assert_close(pred.detach().cpu(), g["pred"], 1e-5, "pred")
    gmax = max(float(np.abs(g["grad." + k]).max()) for k, _ in G.named_parameters())
    for k, p in G.named_parameters():
        if k.endswith(ZERO_GRAD):
            continue
        assert_close(sub(p.grad), g["grad." + k], 1e-5, "grad " + k, atol=2.0 ** -22 * gmax)
    del inter


@pytest.mark.parametrize("train", [False, True])
def test_esat_packed_ragged_bags_vs_oracle(train):
    """Several bags of different lengths in one packed call (attention and GAPool are segmented by bag): per-bag outputs,
    H, the encoder output and the gradients of sum(pred) against the oracle run bag by bag; with positional embedding."""
    C, d = 1024, 384
    Ns = [640, 16, 2064, 160]
    sd = O.synth_state_dict(O.G_ESAT_SHAPES(C, d), 41)
    xs = [O.synth_bag(n, 50 + i, C) for i, n in enumerate(Ns)]
    rng = np.random.default_rng(42)
    noise = torch.tensor(rng.uniform(size=(len(Ns), d // 2)), dtype=torch.float32)
    coords = [torch.tensor(rng.integers(0, 80, size=(n // 16, 2)), dtype=torch.int64) for n in Ns]
    masks = [esat_masks(n // 16, d, 60 + 10 * i) for i, n in enumerate(Ns)] if train else None
    sd, xs = condition_esat_case(sd, xs, coords, [[masks[i] if train else None] for i in range(len(Ns))])
    G = build_G((C, d, d), mode="patch").train(train)
    G.load_state_dict(sd)
    if train:
        G._inject_masks = {k: torch.cat([m[k] for m in masks]).to(torch.uint8).contiguous().cuda() for k in ("sa", "ff1", "ff2", "ga", "gs", "mlp0")}
        G._inject_masks["attn"] = [m["attn"].to(torch.uint8).cuda() for m in masks]
    bags = ops.PackedBags.from_list([x.cuda() for x in xs])
    pred = G.forward_packed(bags, noise=[None, noise.cuda()], coord=torch.cat(coords).cuda())
    pred.sum().backward()
    # reference: the oracle in float64, on a case conditioned away from its ReLU boundaries (tests/util.py)
    sdr = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    want = []
    for i, x in enumerate(xs):
        o = O.generator_forward(sdr, x.double(), [None, noise[i:i + 1].double()], (0, 1), masks[i] if train else None, backbone="patch",
                                coord=coords[i])
        want.append(o["pred"].reshape(-1))
    want = torch.cat(want)
    want.sum().backward()
    assert_close(pred.detach().cpu().reshape(-1), want.detach(), 1e-5, "pred")
    gmax = max(float(v.grad.abs().max()) for v in sdr.values())
    for k, p in G.named_parameters():
        if not k.endswith(ZERO_GRAD):
            assert_close(p.grad.cpu(), sdr[k].grad, 1e-5, "grad " + k, atol=2.0 ** -22 * gmax)


@pytest.mark.parametrize("name", ["g_esat_eval_full", "g_esat_train_full"])
@pytest.mark.parametrize("mode", ["tf32", "bf16"])
def test_esat_reduced_precision_vs_golden(mode, name):
    """tf32 / bf16 modes: the N-row projection of the patch embedding runs on the tcgen05 engine (bf16: x and the
    pre-LayerNorm projection stored in bf16), the region-level contractions on kind::tf32 and the attention on warp-level
    tf32 tensor-core kernels (train fixture: injected masks incl. the attention probabilities): 2e-2 against the fp32 fixture."""
    g = golden(name)
    C, d, N, train, seed, with_coord = [int(v) for v in g["cfg"]]
    G = build_G((C, d, d), mode="patch").train(bool(train))
    G.load_state_dict(O.synth_state_dict(O.G_ESAT_SHAPES(C, d), seed))
    x = O.synth_bag(N, seed, C)
    noise = torch.tensor(np.random.default_rng(seed + 7).uniform(size=(1, d // 2)), dtype=torch.float32)
    G.draw_noise = lambda nb, dev, zero: [None, noise.to(dev)]
    if train:
        G._inject_masks = _dev_masks(esat_masks(N // 16, d, seed * 10))
    advmil_b200.set_precision(mode)
    try:
        pred = G(x.cuda().unsqueeze(0), torch.tensor(g["coord"]).cuda().unsqueeze(0) if with_coord else None)
        pred.sum().backward()
    finally:
        advmil_b200.set_precision("fp32")
    assert_close(pred.detach().cpu(), g["pred"], 2e-2, "pred")
    gmax = max(float(np.abs(g["grad." + k]).max()) for k, _ in G.named_parameters())
    for k, p in G.named_parameters():
        if not k.endswith(ZERO_GRAD):
            assert_close(sub(p.grad), g["grad." + k], 2e-2, "grad " + k, atol=2e-3 * gmax)


@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_esat_in_kernel_dropout_equals_injected_masks_of_the_same_bits(monkeypatch, mode):
    """Train mode with the counter-based in-kernel generator == the oracle fed with the masks advmil_dropout_mask
    materialises for the same seed, including the dropout on the attention probabilities (row = region * nhead + head,
    col = key region); forward and backward regenerate identical bits."""
    from advmil_b200.model import GANSurv
    C, d, N, nh = 1024, 384, 1600, 8
    R = N // 16
    sd = O.synth_state_dict(O.G_ESAT_SHAPES(C, d), 71)
    G = build_G((C, d, d), mode="patch").train()
    G.load_state_dict(sd)
    monkeypatch.setattr(GANSurv, "next_dropout_seed", lambda: 0xABCDEF)
    x = O.synth_bag(N, 72, C)
    noise = torch.tensor(np.random.default_rng(73).uniform(size=(1, d // 2)), dtype=torch.float32)
    advmil_b200.set_precision(mode)
    try:
        pred = G.forward_packed(ops.PackedBags.from_single(x.cuda()), noise=[None, noise.cuda()])
        pred.sum().backward()
    finally:
        advmil_b200.set_precision("fp32")
    tol = 1e-5 if mode == "fp32" else 2e-2
    m = {k: ops.dropout_mask(0xABCDEF, k, p, r, w).cpu().float()
         for k, p, r, w in (("sa", .25, R, d), ("ff1", .25, R, d), ("ff2", .25, R, d), ("ga", .25, R, d), ("gs", .25, R, d),
                            ("mlp0", .6, 1, d // 2))}
    m["attn"] = ops.dropout_mask(0xABCDEF, "attn", .25, R * nh, R).cpu().float().reshape(R, nh, R).permute(1, 0, 2).contiguous()
    assert abs(float(m["attn"].mean()) - 0.75) < 5e-3
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o = O.generator_forward(sdr, x, [None, noise], (0, 1), m, backbone="patch")
    o["pred"].sum().backward()
    assert_close(pred.detach().cpu().reshape(-1), o["pred"].detach().reshape(-1), tol, "pred")
    gmax = max(float(v.grad.abs().max()) for v in sdr.values())
    for k, p in G.named_parameters():
        if not k.endswith(ZERO_GRAD):
            assert_close(p.grad.cpu(), sdr[k].grad, tol, "grad " + k, atol=(2.0 ** -22 if mode == "fp32" else 2e-3) * gmax)


def test_esat_module_surface():
    """load_backbone('patch', dims) mirrors the reference: same state_dict names/shapes, backbone-only forward -> [1, d],
    the handler's placeholder x_ext (dataset/PatchWSI.py:83) fails like the reference, 'graph' stays out of scope."""
    from advmil_b200.model.backbone import load_backbone
    bb = load_backbone("patch", [1024, 384, 384]).cuda().eval()
    want = {k[len("backbone."):]: v for k, v in O.G_ESAT_SHAPES().items() if k.startswith("backbone.")}
    assert {k: tuple(v.shape) for k, v in bb.state_dict().items()} == want
    x = torch.randn(1, 320, 1024, device="cuda")
    with torch.no_grad():
        assert bb(x, None).shape == (1, 384)
    with pytest.raises(IndexError):
        bb(x, torch.Tensor([0]).unsqueeze(0).cuda())
    with pytest.raises(AssertionError):
        bb(torch.randn(1, 100, 1024, device="cuda"), None)
    with pytest.raises(NotImplementedError):
        load_backbone("graph", [1024, 384, 384])
