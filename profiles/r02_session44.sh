cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_step.py tests/test_gpu_esat.py -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -6 | cut -c1-300
timeout 600 python bench.py --backbone patch --no-cpu-baseline > gpurun_out/bench_r02_esat.json 2> gpurun_out/bench_r02_esat.err; echo "esat bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_esat.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('gpu_launches'), 'e2e', d['e2e']['value'], d['roofline']['achieved'])
PY
