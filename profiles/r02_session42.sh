cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python profiles/esat_bench.py --modes bf16 --steps 10 2>&1 | grep "^{" | tee gpurun_out/r02_esat_generator_bench.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in d if k in ('what','ms_per_call','ms_per_step','bags_per_s')})"
timeout 600 python bench.py --backbone patch --no-cpu-baseline > gpurun_out/bench_r02_esat.json 2> gpurun_out/bench_r02_esat.err; echo "esat bench rc=$?"; tail -2 gpurun_out/bench_r02_esat.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_esat.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('gpu_launches'), 'e2e', d['e2e']['value'], d['config'].get('engine'))
PY
