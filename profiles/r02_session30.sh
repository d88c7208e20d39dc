cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -6 | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_s30.json 2> gpurun_out/bench_s30.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s30.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
print({k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in d['configs'].items()})
print({k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in d['modes'].items()})
print('dropin', {k:(round(v['value']),round(v['ms_per_step'],1)) for k,v in d['dropin_handler'].items() if isinstance(v,dict)}, 'eager', round(d['gpu_eager_baseline']['value']), 'cpu', d['cpu_baseline']['value'])
PY
timeout 300 python profiles/esat_bench.py --modes bf16 --steps 10 2>&1 | grep "^{" | cut -c1-160
