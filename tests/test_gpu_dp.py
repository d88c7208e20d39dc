"""GPU data parallel: a 2-rank NCCL AdvStep on sharded bags updates the parameters exactly like one rank on all bags
(SURVEY.md §8e correctness check).  Needs >= 2 GPUs (skipped otherwise): run with `gpurun --gpus 2`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

NS = [320, 1600, 160, 640, 96, 2048]


def _build(dev):
    from oracle import advmil_oracle as O
    from tests.util import build_D, build_G
    G, D = build_G(device=dev), build_D(device=dev)
    G.load_state_dict(O.synth_state_dict(O.G_SHAPES(), 1))
    D.load_state_dict(O.synth_state_dict(O.D_SHAPES(), 2))
    G.backbone.p = 0.0          # dropout off so that sharded and single-process runs see identical arithmetic
    G.p_head = 0.0
    D.net_pair_one.p = 0.0
    return G, D


def _inputs():
    from oracle import advmil_oracle as O
    xs = [O.synth_bag(n, 80 + i) for i, n in enumerate(NS)]
    t, e = O.synth_labels(len(NS), 5)
    e[0] = 1.0
    vis = torch.tensor([1, 1, 0, 1, 1, 1], dtype=torch.uint8)
    rng = np.random.default_rng(6)
    nd = torch.tensor(rng.uniform(size=(len(NS), 192)), dtype=torch.float32)
    ng = torch.tensor(rng.uniform(size=(len(NS), 192)), dtype=torch.float32)
    return xs, t, e, vis, nd, ng


def _run(rank, world, port, q):
    from advmil_b200 import ops
    from advmil_b200.dataset.packed import shard_bags_balanced
    from advmil_b200.step import AdvStep
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                             device_id=dev)
    G, D = _build(dev)
    eng = AdvStep(G, D)
    xs, t, e, vis, nd, ng = _inputs()
    mine = shard_bags_balanced(NS, world)[rank]
    bags = ops.PackedBags.from_list([xs[i].to(dev) for i in mine])
    out = eng.step(bags, t[mine].to(dev), e[mine].to(dev), vis[mine].to(dev), noise_d=nd[mine].to(dev),
                   noise_g=ng[mine].to(dev))
    torch.cuda.synchronize()
    losses = out["losses"].clone()
    if world > 1:
        torch.distributed.all_reduce(losses[:4])          # loss kernels write rank-local contributions
    if rank == 0:
        q.put((eng.G.grad.cpu().numpy(), eng.D.grad.cpu().numpy(), losses.cpu().numpy()))
    if world > 1:
        torch.distributed.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p

def _launch(target, world, timeout=240):
    """Runs `target(rank, world, port, q)` on `world` spawned ranks and returns what rank 0 put on the queue.  Never blocks
    for ever: a rank that dies (or the deadline) fails the test and the surviving ranks are terminated."""
    import queue as _queue
    import time
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out, t0 = None, time.time()
    try:
        while out is None:
            try:
                out = q.get(timeout=2.0)
            except _queue.Empty:
                dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
                assert not dead, f"a rank exited with {dead} before rank 0 reported"
                assert time.time() - t0 < timeout, "ranks did not report in time"
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
    finally:
        for p in procs:
            if p.is_alive():
                p.terminate()
    return out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_equals_single_rank():
    res = {world: _launch(_run, world) for world in (1, 2)}
    g1, d1, l1 = res[1]
    g2, d2, l2 = res[2]
    # all-reduced gradients of both networks and the global losses; only the fp32 reduction order differs
    assert float(np.abs(g1 - g2).max()) <= 1e-5 * float(np.abs(g1).max())
    assert float(np.abs(d1 - d2).max()) <= 1e-5 * float(np.abs(d1).max())
    assert float(np.abs(l1[:4] - l2[:4]).max()) < 1e-5


def _run_esat(rank, world, port, q):
    """ModuleAdvStep with the ESAT generator, eval-style arithmetic (dropout probabilities zeroed) on sharded bags."""
    from advmil_b200 import ops
    from advmil_b200.dataset.packed import shard_bags_balanced
    from advmil_b200.step import ModuleAdvStep
    from oracle import advmil_oracle as O
    from tests.util import build_D, build_G
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                             device_id=dev)
    G, D = build_G((1024, 384, 384), mode="patch", device=dev), build_D(device=dev)
    G.load_state_dict(O.synth_state_dict(O.G_ESAT_SHAPES(), 3))
    D.load_state_dict(O.synth_state_dict(O.D_SHAPES(), 4))
    G.backbone.p = G.backbone.pool.p = 0.0
    G.p_head = 0.0
    D.net_pair_one.p = 0.0
    eng = ModuleAdvStep(G, D)
    xs, t, e, vis, nd, ng = _inputs()
    mine = shard_bags_balanced(NS, world)[rank]
    bags = ops.PackedBags.from_list([xs[i].to(dev) for i in mine])
    out = eng.step(bags, t[mine].to(dev), e[mine].to(dev), vis[mine].to(dev), noise_d=nd[mine].to(dev), noise_g=ng[mine].to(dev))
    torch.cuda.synchronize()
    losses = torch.stack([out["dis_loss"], out["gen_loss"], out["t_reg_loss"]])
    if world > 1:
        torch.distributed.all_reduce(losses)              # every rank holds its bags' share of the globally normalised losses
    if rank == 0:
        q.put((eng.G.grad.cpu().numpy(), eng.D.grad.cpu().numpy(), losses.cpu().numpy()))
    if world > 1:
        torch.distributed.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_module_step_with_esat_equals_single_rank():
    """The same check for ModuleAdvStep (ESAT generator): global-count loss normalisation + flat-buffer all-reduce give the
    single-process gradients of both networks."""
    res = {world: _launch(_run_esat, world) for world in (1, 2)}
    g1, d1, l1 = res[1]
    g2, d2, l2 = res[2]
    assert float(np.abs(g1 - g2).max()) <= 1e-5 * float(np.abs(g1).max())
    assert float(np.abs(d1 - d2).max()) <= 1e-5 * float(np.abs(d1).max())
    assert float(np.abs(l1 - l2).max()) < 1e-5
