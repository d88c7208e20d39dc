cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -6 | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_final_ref.json 2>/dev/null; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['h2d_gbs_per_gpu'], 'launches', d['gpu_launches'])
print('roofline', d['roofline'])
print({k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in d['configs'].items()})
print({k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in d['modes'].items()})
print('sustained', round(d['sustained']['value']), d['sustained']['clocks']['sm_mhz'])
print('dropin', {k:(round(v['value']),round(v['ms_per_step'],1)) for k,v in d['dropin_handler'].items() if isinstance(v,dict)}, 'eager', round(d['gpu_eager_baseline']['value'],1), 'cpu', round(d['cpu_baseline']['value'],2))
r=json.loads(open('gpurun_out/bench_r02_final_ref.json').read().strip().splitlines()[-1]); print('ref arm', r['value'])
PY
