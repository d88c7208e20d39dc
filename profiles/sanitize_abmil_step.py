import os, sys, torch
sys.path.insert(0, os.getcwd())
from advmil_b200 import ops
from advmil_b200.step import AdvStep, sample_inference
from tests.util import build_D, build_G
mode = sys.argv[1]
Ns = [640, 16, 2064, 160]
torch.manual_seed(0)
G, D = build_G(), build_D()
eng = AdvStep(G, D, precision=mode)
x = torch.randn(sum(Ns), 1024, device="cuda")
bags = ops.PackedBags(x.to(torch.bfloat16) if mode == "bf16" else x, Ns)
nb = len(Ns)
t = torch.rand(nb, device="cuda"); e = torch.ones(nb, device="cuda"); vis = torch.ones(nb, dtype=torch.uint8, device="cuda")
for i in range(2):
    out = eng.step(bags, t, e, vis)
print(mode, eng.loss_dict(out))
G.eval(); D.eval()
r = sample_inference(G, D, ops.PackedBags(x, Ns), 30)
torch.cuda.synchronize()
print("sample", r["avg_y_hat"].reshape(-1).tolist())
