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
