"""CPU: the torch.library operators on the meta device -- schema, fake implementations and the autograd registration
(forward shapes, a backward pass that reaches every parameter, absent parameters and injected masks as placeholders).
No kernel runs; the CUDA implementations are checked in tests/test_gpu_library.py."""
import pytest
import torch

from advmil_b200 import library, ops
from tests.util import build_D, build_G


def _meta(ps):
    return [None if p is None else torch.empty(p.shape, dtype=p.dtype, device="meta", requires_grad=p.requires_grad) for p in ps]


def _bags(lengths, C=1024, dtype=torch.float32):
    b = ops.PackedBags.__new__(ops.PackedBags)
    rows = sum(lengths)
    b.x = torch.empty(rows, C, device="meta", dtype=dtype)
    b.lengths, b.rows, b.C, b.bags = list(lengths), rows, C, len(lengths)
    b.offsets = torch.empty(len(lengths) + 1, dtype=torch.int32, device="meta")
    return b


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "bf16"])
@pytest.mark.parametrize("masks", [None, "one", "all"])
def test_generator_and_discriminator_ops_differentiate_on_meta(precision, masks):
    G, D = build_G(device="cpu"), build_D(device="cpu")
    G.train(); D.train()
    gp, dp = _meta(G.gen_params()), _meta(D.disc_params())
    bags = _bags([320, 640])
    noise = torch.empty(2, 192, device="meta")
    prec = ops.PRECISIONS[precision]
    u8 = dict(dtype=torch.uint8, device="meta")
    gm = dm = None
    if masks == "one":
        gm, dm = {"h": torch.empty(960, 512, **u8)}, {"fc2": torch.empty(1, 64, **u8)}
    elif masks == "all":
        gm = {"h": torch.empty(960, 512, **u8), "a": torch.empty(960, 384, **u8), "b": torch.empty(960, 384, **u8),
              "rho": torch.empty(1, 384, **u8), "mlp0": torch.empty(1, 192, **u8)}
        dm = {"fc1": torch.empty(60, 64, **u8), "ga": torch.empty(60, 128, **u8), "gs": torch.empty(60, 128, **u8),
              "fc2": torch.empty(1, 64, **u8)}
    pred = library.generator(G.config(), bags, None, noise, True, 11, gm, prec, gp)
    f = library.discriminator(D.config(), bags, pred, True, 12, dm, prec, dp)
    assert tuple(pred.shape) == (2,) and tuple(f.shape) == (2,) and pred.dtype == torch.float32
    (f.sum() + pred.sum()).backward()
    for p in gp + dp:
        assert p.grad is not None and p.grad.shape == p.shape


def test_absent_parameters_travel_as_placeholders():
    """Backbone-only generator (no head tensors) and a bag-level discriminator: None entries in the C ABI's tensor order."""
    G = build_G(device="cpu")
    gp = _meta(G.gen_params())
    gp[10:] = [None] * 4                                   # no noise-MLP head: the op returns H [bags, o]
    bags = _bags([320])
    cfg = G.config()
    H = library.generator(cfg, bags, None, None, False, 0, None, ops.FP32, gp)
    assert tuple(H.shape) == (1, cfg.o)
    H.sum().backward()
    assert all(p.grad is not None for p in gp[:10] if p is not None)
    # rows that require grad (the DeepAttMISL path feeds cluster means): dx comes back
    x = torch.empty(320, 1024, device="meta", requires_grad=True)
    H = library.generator(cfg, bags, None, None, False, 0, None, ops.FP32, _meta(G.gen_params())[:10] + [None] * 4, x_grad=x)
    H.sum().backward()
    assert x.grad is not None and x.grad.shape == x.shape


def test_discriminator_t_gradient_and_frozen_parameters():
    """The G step: D's parameters frozen (requires_grad False), the gradient flows to the prediction t only."""
    D = build_D(device="cpu")
    for p in D.parameters():
        p.requires_grad_(False)
    dp = _meta(D.disc_params())
    bags = _bags([320, 640], dtype=torch.bfloat16)
    t = torch.empty(2, device="meta", requires_grad=True)
    f = library.discriminator(D.config(), bags, t, False, 0, None, ops.PRECISIONS["bf16"], dp)
    f.sum().backward()
    assert t.grad is not None and t.grad.shape == t.shape
    assert all(p.grad is None for p in dp if p is not None)


def test_module_surface_traces_on_meta_through_the_registered_ops():
    """The drop-in modules themselves (reference signatures: G(x [1,N,C], x_ext), D(x, t)) dispatch to the registered operators
    when they are traced: on the meta device a forward + backward runs shape-only, without any kernel."""
    G, D = build_G(device="meta"), build_D(device="meta")
    G.train(); D.train()
    x = torch.empty(1, 640, 1024, device="meta")
    pred = G(x, None)
    assert tuple(pred.shape) == (1, 1)
    f = D(x, pred)
    assert tuple(f.shape) == (1, 1)
    (f.sum() + pred.sum()).backward()
    assert all(p.grad is not None and p.grad.shape == p.shape for p in list(G.parameters()) + list(D.parameters()))
