"""Tabulates gpurun_out/r02_bf16_parity.json (written by tests/test_gpu_bf16_step.py on the GPU box) into
profiles/r02_bf16_parity.md: per gradient tensor of the C-fused bf16 AdvStep, the worst error over the four cases
(ragged / 16 x 16384, randn / non-negative features), step 0, against the fp32 oracle and the bf16-storage oracle.

    python profiles/bf16_parity_table.py [json] [md]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_bf16_parity.json")
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02_bf16_parity.md")
d = json.load(open(src))
cases = list(d.keys())
tensors = {}
for case, rows in d.items():
    for k, v in rows.items():
        if not isinstance(v, dict) or not k.startswith("step0."):
            continue
        name = k[len("step0."):]
        t = tensors.setdefault(name, {})
        t[case] = v
lines = ["# bf16 mode: gradient parity of the C-fused AdvStep (round 2)", "",
         "Source: `tests/test_gpu_bf16_step.py` on a B200 (`gpurun_out/r02_bf16_parity.json`, copied to `profiles/r02_bf16_parity.json`).",
         "One D step + G step (`model/model_handler.py:349-498`) with injected dropout masks; errors are max |cuda - oracle| over the",
         "tensor divided by max |oracle| of the tensor (`max`) and the relative L2 error (`l2`).  `fp32` = the oracle in the reference's",
         "fp32 arithmetic; `emu` = the oracle emulating the bf16 storage points of the CUDA path (`oracle.bf16_storage()`).",
         "Cases: " + ", ".join(f"`{c}`" for c in cases) + ".  Worst case over the cases is shown per tensor.", "",
         "| net.tensor | max vs fp32 | l2 vs fp32 | worst case | max vs emu | l2 vs emu | within 2e-2 of fp32 (max) |",
         "|---|---:|---:|---|---:|---:|---|"]
n_ok = 0
for name in sorted(tensors):
    t = tensors[name]
    wc = max(t, key=lambda c: t[c]["rel_fp32"])
    mf = t[wc]["rel_fp32"]
    lf = max(v.get("l2_fp32", float("nan")) for v in t.values())
    me = max(v["rel_emu"] for v in t.values())
    le = max(v.get("l2_emu", float("nan")) for v in t.values())
    ok = mf <= 2e-2
    n_ok += ok
    lines.append(f"| {name} | {mf:.1e} | {lf:.1e} | {wc} | {me:.1e} | {le:.1e} | {'yes' if ok else 'no'} |")
lines += ["", f"{n_ok} of {len(tensors)} gradient tensors are within 2e-2 (max-norm) of the fp32 oracle in every case; every tensor is within "
          f"{max(max(v['rel_emu'] for v in t.values()) for t in tensors.values()):.1e} of the bf16-storage oracle.", ""]
post = {c: {k: v for k, v in rows.items() if k.endswith("post_adam_worst_in_lr")} for c, rows in d.items()}
lines += ["Post-Adam parameters (entries whose reference gradient is at least half of the tensor's largest; difference of the parameter",
          "update against the fp32 oracle trainer, in units of the learning rate):", ""]
for c, rows in post.items():
    lines.append(f"* `{c}`: " + ", ".join(f"{k.replace('.post_adam_worst_in_lr', '')} {v:.1e} lr" for k, v in sorted(rows.items())))
open(dst, "w").write("\n".join(lines) + "\n")
json.dump(d, open(os.path.join(ROOT, "profiles", "r02_bf16_parity.json"), "w"), indent=1)
print(dst, len(tensors), "tensors;", n_ok, "within 2e-2 of fp32")
