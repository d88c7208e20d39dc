"""Shared helpers for the tests: build advmil_b200 modules / oracle state dicts from the same synthetic parameters."""
import os
from types import SimpleNamespace

import numpy as np
import torch

from oracle import advmil_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def sub(v, n=512):
    v = v.detach().reshape(-1).double().cpu().numpy()
    stride = max(1, v.size // n)
    return v[::stride].copy()


def build_G(dims=(1024, 384, 384), mode="abmil", gen_dropout=0.6, device="cuda"):
    from advmil_b200.model.backbone import load_backbone
    from advmil_b200.model.GANSurv import Generator
    backbone = load_backbone(mode, list(dims))
    args_noise = SimpleNamespace(noise=[0, 1], hops=1, noise_dist="uniform")
    return Generator(dims[2], 1, backbone, args_noise, False, gen_dropout, "sigmoid").to(device)


def build_D(C=1024, d=128, iprd="instance", prj="x", device="cuda"):
    from advmil_b200.model.GANSurv import PrjDiscriminator
    ax = SimpleNamespace(in_dim=C, out_dim=d, ksize=1, backbone="avgpool", dropout=0.25)
    ay = SimpleNamespace(in_dim=1, hid_dims=[64, 128] if d == 128 else [d // 2, d], norm=False, dropout=0.0)
    return PrjDiscriminator(ax, ay, prj_path=prj, inner_product=iprd).to(device)


def g_masks(N, h, o, seed):
    return {"h": O.synth_masks((N, h), 0.75, seed), "a": O.synth_masks((N, h), 0.75, seed + 1),
            "b": O.synth_masks((N, h), 0.75, seed + 2), "rho": O.synth_masks((1, o), 0.75, seed + 3),
            "mlp0": O.synth_masks((1, o // 2), 0.4, seed + 4)}


def d_masks(R, d, seed):
    return {"fc1": O.synth_masks((R, d // 2), 0.75, seed), "ga": O.synth_masks((R, d), 0.75, seed + 1),
            "gs": O.synth_masks((R, d), 0.75, seed + 2), "fc2": O.synth_masks((1, d // 2), 0.75, seed + 3)}


def esat_masks(R, d, seed, nhead=8):
    return {"attn": O.synth_masks((nhead, R, R), 0.75, seed), "sa": O.synth_masks((R, d), 0.75, seed + 1),
            "ff1": O.synth_masks((R, d), 0.75, seed + 2), "ff2": O.synth_masks((R, d), 0.75, seed + 3),
            "ga": O.synth_masks((R, d), 0.75, seed + 4), "gs": O.synth_masks((R, d), 0.75, seed + 5),
            "mlp0": O.synth_masks((1, d // 2), 0.4, seed + 6)}


def to_dev_masks(m, device="cuda"):
    return {k: v.to(torch.uint8).contiguous().to(device) for k, v in m.items()}


def assert_close(actual, expected, rtol, name="", atol_scale=None, atol=0.0):
    """|a - e| <= rtol * |e| + rtol * scale + atol, scale = max|e| of the tensor (norm-wise floor for entries that
    cancel); atol is an absolute floor used only for gradient tensors (see grad_floor)."""
    a = torch.as_tensor(np.asarray(actual), dtype=torch.float64).reshape(-1)
    e = torch.as_tensor(np.asarray(expected), dtype=torch.float64).reshape(-1)
    assert a.shape == e.shape, f"{name}: shape {a.shape} vs {e.shape}"
    scale = float(e.abs().max()) if atol_scale is None else atol_scale
    err = (a - e).abs()
    tol = rtol * e.abs() + rtol * scale + atol + 1e-12
    bad = err > tol
    assert not bool(bad.any()), (f"{name}: {int(bad.sum())}/{a.numel()} outside rtol={rtol}; max err {float(err.max()):.3e} "
                                 f"(scale {scale:.3e}, rel-to-scale {float(err.max()) / (scale + 1e-30):.3e})")


def grad_floor(ref_grads) -> float:
    """Absolute floor for gradient comparisons: one fp32 ulp (2^-23) of the largest gradient entry of the whole network.
    Gradients that pass through the softmax Jacobian w_n (g.v_n - g.z) of a nearly-uniform attention, or that sum real
    and fake pair contributions of opposite sign, are differences of nearly equal numbers; their fp32 rounding error is
    set by the size of the cancelled terms, not of the result, so two correct fp32 evaluations (e.g. the reference on
    CPU and on GPU) differ there by more than 1e-5 of the tiny result."""
    m = 0.0
    for g in ref_grads:
        if g is not None:
            m = max(m, float(np.abs(np.asarray(g)).max()))
    return 2.0 ** -23 * m


def condition_esat_case(sd, xs, coords, mask_variants, margin=2e-5, iters=30):
    """Moves a synthetic ESAT case away from its ReLU boundaries.  A few of the ~10^6 ReLU arguments of a case (LayerNorm
    outputs of the patch embedding, linear1 and MLPs.0 pre-activations) always land within ~1e-6 of zero, where two
    correct fp32 evaluations (torch's CPU kernels and the CUDA path, which sum in different orders) may take different
    branches; one flipped element moves the affected gradients by 1e-3 of their scale.  Rows of x whose embedding has
    such an element get a small deterministic perturbation, offending linear1 / MLPs.0 columns a bias nudge, until the
    float64 oracle sees no argument closer than `margin` to zero (for every mask variant)."""
    import torch.nn.functional as F
    sd = {k: v.clone() for k, v in sd.items()}
    xs = [x.clone() for x in xs]
    P, L = "backbone.patch_embedding_layer.", "backbone.patch_encoder_layer.layers.0."
    g = torch.Generator().manual_seed(12345)
    for _ in range(iters):
        s64 = {k: v.double() for k, v in sd.items()}
        d = s64[P + "conv.bias"].shape[0]
        Wc = s64[P + "conv.weight"].reshape(d, -1)
        clean = True
        for i, x in enumerate(xs):
            ln = F.layer_norm(F.linear(x.double(), Wc, s64[P + "conv.bias"]), (d,), s64[P + "norm.weight"], s64[P + "norm.bias"], 1e-5)
            bad = torch.nonzero((ln.abs() < margin).any(dim=1)).reshape(-1)
            if bad.numel():
                xs[i][bad] += 1e-2 * torch.randn(bad.numel(), x.shape[1], generator=g)
                clean = False
        if not clean:
            continue
        for i, x in enumerate(xs):
            for masks in mask_variants[i]:
                o = O.esat_forward(s64, x.double(), None if coords is None else coords[i], masks)
                pre = F.linear(o["x1"], s64[L + "linear1.weight"], s64[L + "linear1.bias"])
                cols = torch.nonzero((pre.abs() < margin).any(dim=0)).reshape(-1)
                if cols.numel():
                    sd[L + "linear1.bias"][cols] += 1e-3
                    clean = False
                hp = F.linear(o["H"], s64["MLPs.0.0.weight"], s64["MLPs.0.0.bias"])
                cols = torch.nonzero((hp.abs() < margin).any(dim=0)).reshape(-1)
                if cols.numel():
                    sd["MLPs.0.0.bias"][cols] += 1e-3
                    clean = False
        if clean:
            return sd, xs
    raise AssertionError("could not condition the ESAT case")
