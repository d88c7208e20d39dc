"""compute-sanitizer target for the kernels added in round 2: the C-fused ESAT step (tcgen05 attention forward / backward,
region-chain kernel, side stream of the backward), and the vl transport decoder through the feeder.
    compute-sanitizer --tool memcheck python profiles/sanitize_r02_new_kernels.py"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from types import SimpleNamespace as NS
from advmil_b200 import ops
from advmil_b200.dataset.packed import DeviceFeeder, pack_step
from advmil_b200.model.backbone import load_backbone
from advmil_b200.model.GANSurv import Generator, PrjDiscriminator
from advmil_b200.step import EsatAdvStep
Ns = [640, 16, 2064, 160, 4800]
torch.manual_seed(0)
G = Generator(384, 1, load_backbone("patch", [1024, 384, 384]), NS(noise=[0, 1], hops=1, noise_dist="uniform"), False, 0.6, "sigmoid").cuda()
D = PrjDiscriminator(NS(in_dim=1024, out_dim=128, ksize=1, backbone="avgpool", dropout=0.25), NS(in_dim=1, hid_dims=[64, 128], norm=False, dropout=0.0), prj_path="x", inner_product="instance").cuda()
eng = EsatAdvStep(G, D, precision="bf16")
xs = [torch.randn(n, 1024) for n in Ns]
steps = [pack_step(xs, [(0.1 * i, 1.0) for i in range(len(Ns))], dtype=torch.bfloat16).packvl() for _ in range(2)]
coord = torch.randint(0, 200, (sum(Ns) // 16, 2), device="cuda")
for s in DeviceFeeder(steps, device="cuda", depth=2):
    out = eng.step(s.bags, s.t, s.e, s.visible, coord=coord)
torch.cuda.synchronize()
print(eng.loss_dict(out), "finite:", all(bool(torch.isfinite(p).all()) for p in G.parameters()))
