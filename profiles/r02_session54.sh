cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -6 | cut -c1-300
echo "pytest $(( $(date +%s) - t0 )) s"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench_r02_final2.json 2> gpurun_out/bench_r02_final2.err; echo "bench rc=$? $(( $(date +%s) - t0 )) s"
timeout 600 python bench.py --backbone patch > gpurun_out/bench_r02_esat2.json 2> gpurun_out/bench_r02_esat2.err; echo "esat rc=$? $(( $(date +%s) - t0 )) s"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 1500 --csv --log-file gpurun_out/launches_r02_bf16_v2.csv python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > /dev/null 2>&1; echo "list rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:rlip_chain' -s 6 -c 2 -f -o gpurun_out/r02_chain_mma_final python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_final2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['h2d_gbs_per_gpu'], 'launches', d['gpu_launches'])
print('roofline', d['roofline'])
print({k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in d['configs'].items()})
print({k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in d['modes'].items()})
print('sustained', round(d['sustained']['value']), d['sustained']['clocks']['sm_mhz'])
print('dropin', {k:(round(v['value']),round(v['ms_per_step'],1)) for k,v in d['dropin_handler'].items() if isinstance(v,dict)}, 'eager', round(d['gpu_eager_baseline']['value'],1), 'cpu', round(d['cpu_baseline']['value'],2))
print({k:round(v['ms_per_launch']*1000,1) for k,v in d['kernels'].items()})
e=json.loads(open('gpurun_out/bench_r02_esat2.json').read().strip().splitlines()[-1])
print('esat', e['value'], e['ms_per_step'], 'e2e', e['e2e']['value'])
PY
echo "total $(( $(date +%s) - t0 )) s"
