cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "cluster" 2>&1 | grep -v "^  \|Warning\|^$" | tail -6 | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_s29.json 2> gpurun_out/bench_s29.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s29.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
print(d['configs'])
PY
