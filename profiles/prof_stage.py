"""Stand-alone driver for ncu captures of single hot kernels at the benchmark size (262,144 rows):
    python profiles/prof_stage.py gate_train|gate_eval|pool_bwd|proj|embed [precision]
Runs the stage 3 times (ncu: -k regex:<kernel> -s 2 -c 1 captures the warm third launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from advmil_b200 import ops

what = sys.argv[1]
prec = ops.PRECISIONS[sys.argv[2] if len(sys.argv) > 2 else "bf16"]
rows, nb = 262144, 16
g = torch.Generator(device="cuda").manual_seed(0)
dt = ops.act_dtype(prec)
if what in ("gate_train", "gate_eval", "pool_bwd"):
    v = torch.randn(rows, 384, device="cuda", generator=g).to(dt)
    Wa, Wb = [torch.randn(384, 384, device="cuda", generator=g) / 20 for _ in range(2)]
    ba, bb, wc = [torch.randn(384, device="cuda", generator=g) * 0.1 for _ in range(3)]
    bc = torch.zeros(1, device="cuda")
    for _ in range(3):
        s, ab = ops.gated_score_forward(v, Wa, ba, Wb, bb, wc, bc, p_drop=0.25, seed=7, train=(what != "gate_eval"),
                                        precision=prec, save=(what != "gate_eval"))
elif what == "proj":
    x = torch.randn(rows, 1024, device="cuda", generator=g).to(dt)
    W = torch.randn(384, 1024, device="cuda", generator=g) / 32
    b = torch.zeros(384, device="cuda")
    for _ in range(3):
        ops.linear_forward(x, W, b, act=1, precision=prec)
torch.cuda.synchronize()
print("done", what)
