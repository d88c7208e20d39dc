cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r02_bf16_parity.json
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -50 | cut -c1-300
timeout 300 python bench.py --no-extra-legs --no-cpu-baseline --steps 20 > gpurun_out/bench_s5.json 2>/dev/null; echo "bench rc=$?"
timeout 300 python bench.py --no-extra-legs --no-cpu-baseline --steps 10 --precision tf32x3 > gpurun_out/bench_s5_x3.json 2>/dev/null; echo "bench x3 rc=$?"
python - <<'PY'
import json
for f in ('bench_s5','bench_s5_x3'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
    for k,v in d['kernels'].items(): print('   ',k,{a:round(b,4) for a,b in v.items()})
PY
