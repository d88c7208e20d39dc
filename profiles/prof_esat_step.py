"""Host-side profile of ModuleAdvStep (ESAT generator): wall time per step without sync vs device time."""
import cProfile, pstats, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace as NS
import torch
import advmil_b200
from advmil_b200 import ops
from advmil_b200.model.backbone import load_backbone
from advmil_b200.model.GANSurv import Generator, PrjDiscriminator
from advmil_b200.step import ModuleAdvStep
torch.manual_seed(0)
bags_n, rows = 16, 16384
G2 = Generator(384, 1, load_backbone("patch", [1024, 384, 384]), NS(noise=[0, 1], hops=1, noise_dist="uniform"), False, 0.6, "sigmoid").cuda()
D2 = PrjDiscriminator(NS(in_dim=1024, out_dim=128, ksize=1, backbone="avgpool", dropout=0.25),
                      NS(in_dim=1, hid_dims=[64, 128], norm=False, dropout=0.0), prj_path="x", inner_product="instance").cuda()
eng = ModuleAdvStep(G2, D2, precision="bf16")
x = torch.randn(bags_n * rows, 1024, device="cuda").to(torch.bfloat16)
bags = ops.PackedBags(x, [rows] * bags_n)
t = torch.rand(bags_n, device="cuda"); e = (torch.rand(bags_n, device="cuda") < 0.35).float(); e[0] = 1.0
vis = torch.ones(bags_n, dtype=torch.uint8, device="cuda"); nz = torch.rand(bags_n, 192, device="cuda")
counts = (float(e.sum()), float(bags_n), float(bags_n))
def step(): return eng.step(bags, t, e, vis, noise_d=nz, noise_g=nz, global_counts=counts)
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue {1e3*(t1-t0)/20:.3f} ms/step, with drain {1e3*(t2-t0)/20:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(40); print(s.getvalue()[:7000])
