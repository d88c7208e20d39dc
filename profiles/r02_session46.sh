cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_bf16_step.py tests/test_gpu_fullsize.py tests/test_gpu_parity_modules.py tests/test_gpu_tf32x3.py -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -6 | cut -c1-300
for ov in 1 0; do
ADVMIL_HEAD_OVERLAP=$ov timeout 300 python bench.py --no-extra-legs --no-cpu-baseline > gpurun_out/bench_s46_$ov.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_s46_$ov.json').read().strip().splitlines()[-1])
print('overlap=$ov', round(d['value']), d['ms_per_step'], 'head_bwd', d['kernels']['head_bwd']['ms_per_launch'], 'e2e', round(d['e2e']['value']))
PY
done
