"""GPU: the torch.library operators (`advmil_b200::generator_fwd/_bwd`, `advmil_b200::discriminator_fwd/_bwd`) that the
drop-in modules call: `torch.library.opcheck` (schema, fake implementation vs real shapes/dtypes, autograd registration),
equality with the autograd.Function glue they replace, and FakeTensor tracing of a module call."""
import numpy as np
import pytest
import torch

import advmil_b200
from advmil_b200 import library, ops
from oracle import advmil_oracle as O
from tests.util import build_D, build_G

pytestmark = pytest.mark.gpu


def _nets():
    G, D = build_G(), build_D()
    G.load_state_dict(O.synth_state_dict(O.G_SHAPES(), 1))
    D.load_state_dict(O.synth_state_dict(O.D_SHAPES(), 2))
    return G, D


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_registered_ops_equal_the_function_glue(precision):
    G, D = _nets()
    xs = [O.synth_bag(n, 3 + i).cuda() for i, n in enumerate([320, 640])]
    bags = ops.PackedBags.from_list(xs)
    noise = torch.rand(2, 192, device="cuda")
    prec = ops.PRECISIONS[precision]
    G.train(); D.train()
    pred = library.generator(G.config(), bags, None, noise, True, 11, None, prec, G.gen_params())
    f = library.discriminator(D.config(), bags, pred, True, 12, None, prec, D.disc_params())
    (f.sum() + pred.sum()).backward()
    got = [p.grad.clone() for p in list(G.parameters()) + list(D.parameters())]
    G.zero_grad(); D.zero_grad()
    pred2 = ops.GeneratorFn.apply(G.config(), bags, None, None, noise, True, 11, None, prec, *G.gen_params())
    f2 = ops.DiscriminatorFn.apply(D.config(), bags, pred2, True, 12, None, prec, *D.disc_params())
    (f2.sum() + pred2.sum()).backward()
    assert torch.equal(pred, pred2) and torch.equal(f, f2)
    for a, p in zip(got, list(G.parameters()) + list(D.parameters())):
        assert torch.equal(a, p.grad)


def test_opcheck_and_fake_tracing():
    G, D = _nets()
    G.eval(); D.eval()
    x = O.synth_bag(320, 5).cuda()
    bags = ops.PackedBags.from_single(x)
    noise = torch.rand(1, 192, device="cuda")
    icfg, fcfg = library.gen_cfg_lists(G.config())
    args = (bags.x, bags.offsets, list(bags.lengths), icfg, fcfg, None, noise, False, 0, ops.FP32, True, library._dense([None] * 5, bags.x), library._dense(G.gen_params(), bags.x))
    torch.library.opcheck(torch.ops.advmil_b200.generator_fwd.default, args, test_utils=("test_schema", "test_faketensor"))
    dcfg, dfcfg = library.disc_cfg_lists(D.config())
    t = torch.rand(1, device="cuda", requires_grad=True)
    dargs = (bags.x, bags.offsets, list(bags.lengths), dcfg, dfcfg, t, False, 0, ops.FP32, True, library._dense([None] * 4, bags.x), library._dense(D.disc_params(), bags.x))
    torch.library.opcheck(torch.ops.advmil_b200.discriminator_fwd.default, dargs, test_utils=("test_schema", "test_faketensor"))
    # FakeTensor tracing of the module surface: shapes without running a kernel
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode(allow_non_fake_inputs=True) as mode:
        fx = mode.from_tensor(x)
        fb = ops.PackedBags.__new__(ops.PackedBags)
        fb.__dict__.update(bags.__dict__)
        fb.x = fx
        out = library.generator(G.config(), fb, None, mode.from_tensor(noise), False, 0, None, ops.FP32, G.gen_params())
        assert tuple(out.shape) == (1,)


def test_activations_are_released_by_reference_counting():
    """The saved activations must not sit in a reference cycle (output -> grad_fn -> ctx -> output): dropping the result of
    a forward frees its activations at once, without the cyclic collector (the drop-in handler runs 16 bags x 5 module
    calls per optimiser step; a cycle made every allocation a cudaMalloc and the step 8x slower)."""
    import gc
    G, D = _nets()
    G.train(); D.train()
    bags = ops.PackedBags.from_single(O.synth_bag(4096, 9).cuda())
    noise = torch.rand(1, 192, device="cuda")

    def fwd():
        pred = library.generator(G.config(), bags, None, noise, True, 3, None, ops.FP32, G.gen_params())
        return library.discriminator(D.config(), bags, pred, True, 4, None, ops.FP32, D.disc_params())

    f = fwd()
    del f
    gc.collect()
    torch.cuda.synchronize()
    base = torch.cuda.memory_allocated()
    gc.disable()
    try:
        f = fwd()
        assert torch.cuda.memory_allocated() > base + 4096 * 512 * 4          # the activations are alive with the graph
        del f
        assert torch.cuda.memory_allocated() <= base + (1 << 20)
    finally:
        gc.enable()
