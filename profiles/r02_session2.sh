cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --deselect "tests/test_gpu_handler.py::test_unmodified_handler_epoch_in_the_fast_modes[tf32x3]" 2>&1 | tail -25
timeout 300 python bench.py --no-extra-legs --no-cpu-baseline --steps 20 > gpurun_out/bench_s2.json 2>/dev/null; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
for k,v in d['kernels'].items(): print('   ',k,{a:round(b,4) for a,b in v.items()})
PY
