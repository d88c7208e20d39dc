"""Decode throughput of the 12-bit transport form (csrc/codec.cu) on one benchmark step: 16 x 16384 x 1024 bf16 elements.
Algorithmic bytes per element: 1.5 read (lo + hi planes) + 2 written.  CUDA events, 20 calls after 3 warm-ups."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from advmil_b200.dataset.codec import decode_p12_device, encode_bf16_p12  # noqa: E402

rows = 16 * 16384
x = torch.randn(rows, 1024).to(torch.bfloat16)
p = encode_bf16_p12(x)
lo, hi, ei, ee = p.lo.cuda(), p.hi.cuda(), p.esc_idx.cuda(), p.esc_exp.cuda()
out = torch.empty(rows, 1024, dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    decode_p12_device(lo, hi, p.table, ei, ee, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    decode_p12_device(lo, hi, p.table, ei, ee, out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
n = rows * 1024
assert torch.equal(out.cpu().view(torch.int16), x.view(torch.int16))
print(json.dumps({"what": "bf16 p12 decode, one 16-bag step", "elements": n, "escapes": int(ei.numel()), "ms": ms,
                  "algorithmic_GBs": 3.5 * n / ms / 1e6, "transport_ratio": p.nbytes / (2 * n)}))
