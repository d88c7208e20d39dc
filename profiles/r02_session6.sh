cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r02_bf16_parity.json
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -40 | cut -c1-300
timeout 300 python bench.py --no-extra-legs --no-cpu-baseline --steps 20 > gpurun_out/bench_s6.json 2>/dev/null; echo "bench rc=$?"
ADVMIL_TC_CLUSTER=2 timeout 300 python bench.py --no-extra-legs --no-cpu-baseline --steps 20 > gpurun_out/bench_s6_cl2.json 2>/dev/null; echo "bench rc=$?"
timeout 300 python bench.py --no-extra-legs --no-cpu-baseline --steps 20 > gpurun_out/bench_s6b.json 2>/dev/null; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('bench_s6','bench_s6_cl2','bench_s6b'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
    print('   ', {k: round(v['ms_per_launch']*1e3,1) for k,v in d['kernels'].items()})
PY
