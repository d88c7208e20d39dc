"""Top stall locations of one kernel of an .ncu-rep (SASS level): python profiles/ncu_hot.py REP LAUNCH_INDEX [TOP]"""
import csv, subprocess, sys, io
rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
print(lines[0][:200])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
h = rows[0]
ci = {n: i for i, n in enumerate(h)}
S, I = ci["# Samples"], ci["Instructions Executed"]
data = [(int(r[S] or 0), n, r) for n, r in enumerate(rows[1:]) if len(r) > S and (r[S] or "0").isdigit()]
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data), "executed", sum(int(d[2][I] or 0) for d in data))
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") or n.lower().startswith("warp stall")]
for s, n, r in sorted(data, reverse=True)[:top]:
    print(f"{s:7d} {100.0*s/max(tot,1):5.1f}%  #{n:5d} exec={r[I]:>8}  {r[ci['Source']].strip()[:110]}")
