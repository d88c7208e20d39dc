import json, sys
d = json.load(open(sys.argv[1]))
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"])
tot = 0
for k, v in d["kernels"].items():
    tot += v["ms_per_launch"] * v["launches"] / d["steps"]
    print(f"  {k:14s} {v['ms_per_launch']*1e3:8.1f} us x{v['launches']:3d} share {v['share_of_step']*100:5.1f}%  ", {a: round(b, 1) for a, b in v.items() if a in ("tflops", "gbs")})
print("tagged ms/step", round(tot, 3), "roofline", d["roofline"])
