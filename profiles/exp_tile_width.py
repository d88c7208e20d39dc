import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from advmil_b200 import ops
rows = 262144
g = torch.Generator(device="cuda").manual_seed(0)
for K in (1024, 384):
    x = torch.randn(rows, K, device="cuda", generator=g).to(torch.bfloat16)
    for N in (128, 384, 512, 768):
        W = torch.randn(N, K, device="cuda", generator=g) / 32
        b = torch.zeros(N, device="cuda")
        for _ in range(2): ops.linear_forward(x, W, b, act=1, precision=ops.BF16)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.linear_forward(x, W, b, act=1, precision=ops.BF16)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"K={K} N={N}: {ms*1e3:.1f} us  {2*rows*K*N/ms/1e9:.0f} TFLOP/s  HBM {(rows*K*2+rows*N*2)/ms/1e6:.0f} GB/s")
