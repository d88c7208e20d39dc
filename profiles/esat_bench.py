"""ESAT generator (bcb_mode 'patch') forward + backward on the benchmark shape: 16 bags x 16384 x 1024 per call, train mode
with in-kernel dropout.  Prints one JSON line per precision mode (CUDA-event timing, warm-up 3, 10 timed calls).
    python profiles/esat_bench.py [--bags 16] [--rows 16384] [--modes bf16,tf32,fp32]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace  # noqa: E402

import advmil_b200  # noqa: E402
from advmil_b200 import ops  # noqa: E402
from advmil_b200.model.backbone import load_backbone  # noqa: E402
from advmil_b200.model.GANSurv import Generator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bags", type=int, default=16)
ap.add_argument("--rows", type=int, default=16384)
ap.add_argument("--modes", default="bf16,tf32,fp32")
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
torch.manual_seed(0)
G = Generator(384, 1, load_backbone("patch", [1024, 384, 384]), SimpleNamespace(noise=[0, 1], hops=1, noise_dist="uniform"), False, 0.6,
              "sigmoid").cuda().train()
x32 = torch.randn(args.bags * args.rows, 1024, device="cuda")
for mode in args.modes.split(","):
    advmil_b200.set_precision(mode)
    x = x32.to(torch.bfloat16) if mode == "bf16" else x32
    bags = ops.PackedBags(x, [args.rows] * args.bags)
    noise = [None, torch.rand(args.bags, 192, device="cuda")]

    def step():
        G.zero_grad(set_to_none=True)
        pred = G.forward_packed(bags, noise=noise)
        pred.sum().backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"what": "ESAT generator fwd+bwd (train mode)", "mode": mode, "bags": args.bags, "rows_per_bag": args.rows,
                      "ms_per_call": ms, "bags_per_s": args.bags / ms * 1e3}))
advmil_b200.set_precision("fp32")

# full adversarial step (D update + G update, both Adam steps) with the ESAT generator and the RLIP discriminator
from types import SimpleNamespace as NS  # noqa: E402

from advmil_b200.model.GANSurv import PrjDiscriminator  # noqa: E402
from advmil_b200.step import EsatAdvStep, ModuleAdvStep  # noqa: E402

for mode, Engine in [(m, E) for m in args.modes.split(",") for E in (ModuleAdvStep, EsatAdvStep)]:
    torch.manual_seed(0)
    G2 = Generator(384, 1, load_backbone("patch", [1024, 384, 384]), SimpleNamespace(noise=[0, 1], hops=1, noise_dist="uniform"), False,
                   0.6, "sigmoid").cuda()
    D2 = PrjDiscriminator(NS(in_dim=1024, out_dim=128, ksize=1, backbone="avgpool", dropout=0.25),
                          NS(in_dim=1, hid_dims=[64, 128], norm=False, dropout=0.0), prj_path="x", inner_product="instance").cuda()
    eng = Engine(G2, D2, precision=mode)
    x = x32.to(torch.bfloat16) if mode == "bf16" else x32
    bags = ops.PackedBags(x, [args.rows] * args.bags)
    t = torch.rand(args.bags, device="cuda")
    e = (torch.rand(args.bags, device="cuda") < 0.35).float()
    e[0] = 1.0
    vis = torch.ones(args.bags, dtype=torch.uint8, device="cuda")
    nz = torch.rand(args.bags, 192, device="cuda")
    counts = (float(e.sum()), float(args.bags), float(args.bags))
    for _ in range(3):
        eng.step(bags, t, e, vis, noise_d=nz, noise_g=nz, global_counts=counts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = eng.step(bags, t, e, vis, noise_d=nz, noise_g=nz, global_counts=counts)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    ld = eng.loss_dict(out) if Engine is EsatAdvStep else {k: float(out[k]) for k in ("dis_loss", "gen_total_loss")}
    print(json.dumps({"what": f"ESAT G + RLIP D adversarial step ({Engine.__name__}: D update + G update + both Adam steps)", "mode": mode,
                      "bags": args.bags, "rows_per_bag": args.rows, "ms_per_step": ms, "bags_per_s": args.bags / ms * 1e3,
                      "dis_loss": ld["dis_loss"], "gen_total_loss": ld["gen_total_loss"]}))
