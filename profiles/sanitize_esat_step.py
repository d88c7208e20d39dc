import os, sys, torch
sys.path.insert(0, os.getcwd())
from types import SimpleNamespace as NS
from advmil_b200 import ops
from advmil_b200.model.backbone import load_backbone
from advmil_b200.model.GANSurv import Generator, PrjDiscriminator
from advmil_b200.step import ModuleAdvStep
mode = sys.argv[1]
Ns = [int(v) for v in sys.argv[2].split(",")]
torch.manual_seed(0)
G = Generator(384, 1, load_backbone("patch", [1024, 384, 384]), NS(noise=[0, 1], hops=1, noise_dist="uniform"), False, 0.6, "sigmoid").cuda()
D = PrjDiscriminator(NS(in_dim=1024, out_dim=128, ksize=1, backbone="avgpool", dropout=0.25), NS(in_dim=1, hid_dims=[64, 128], norm=False, dropout=0.0), prj_path="x", inner_product="instance").cuda()
eng = ModuleAdvStep(G, D, precision=mode)
x = torch.randn(sum(Ns), 1024, device="cuda")
bags = ops.PackedBags(x.to(torch.bfloat16) if mode == "bf16" else x, Ns)
nb = len(Ns)
t = torch.rand(nb, device="cuda"); e = torch.ones(nb, device="cuda"); vis = torch.ones(nb, dtype=torch.uint8, device="cuda")
coord = torch.randint(0, 200, (sum(Ns) // 16, 2), device="cuda")
for i in range(2):
    out = eng.step(bags, t, e, vis, coord=coord)
torch.cuda.synchronize()
print(mode, Ns, {k: float(v) for k, v in out.items() if v is not None and v.numel() == 1}, "finite:", all(bool(torch.isfinite(p).all()) for p in G.parameters()))
