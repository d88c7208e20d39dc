cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r02_final3.json 2> gpurun_out/bench_r02_final3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_final3.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['h2d_gbs_per_gpu'], 'launches', d['gpu_launches'])
print('roofline', d['roofline'])
print({k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in d['configs'].items()})
print({k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in d['modes'].items()})
print('sustained', round(d['sustained']['value']), d['sustained']['clocks']['sm_mhz'])
print('dropin', {k:(round(v['value']),round(v['ms_per_step'],1)) for k,v in d['dropin_handler'].items() if isinstance(v,dict)}, 'eager', round(d['gpu_eager_baseline']['value'],1), 'cpu', round(d['cpu_baseline']['value'],2))
print({k:round(v['ms_per_launch']*1000,1) for k,v in d['kernels'].items()})
PY
