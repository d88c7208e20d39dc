"""Prints a per-kernel table from a bench.py JSON line: python profiles/bench_summary.py <file> (last JSON line is used)."""
import json
import sys

lines = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")]
d = json.loads(lines[-1])
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "launches/step",
      d["gpu_launches"] / d["steps"], "host issue ms/step", round(d.get("host_issue_ms_per_step") or 0, 3),
      "profiled pass ms/step", round(d.get("profiled_pass_ms_per_step") or 0, 3), "clocks", d.get("clocks"))
tot = 0
for k, v in d["kernels"].items():
    n = v.get("launches_per_step", v.get("launches", 0) / d["steps"])
    per_step = v["ms_per_launch"] * n
    tot += per_step
    print(f"  {k:14s} {v['ms_per_launch']*1e3:8.1f} us x{n:4.1f}/step = {per_step*1e3:7.1f} us  share {v['share_of_step']*100:5.1f}%  ",
          {a: round(b, 1) for a, b in v.items() if a in ("tflops", "gbs")})
print("tagged ms/step", round(tot, 3), "roofline", d["roofline"])
