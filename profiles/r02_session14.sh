cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_modules.py tests/test_gpu_step.py tests/test_gpu_bf16.py tests/test_gpu_bf16_step.py tests/test_gpu_fullsize.py tests/test_gpu_tf32x3.py tests/test_gpu_tc.py -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -30 | cut -c1-300
for rt in 4 8; do
ADVMIL_RLIP_CHAIN_RT=$rt timeout 600 python bench.py --no-extra-legs --no-cpu-baseline > gpurun_out/bench_s14_rt$rt.json 2> gpurun_out/bench_s14.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_s14_rt$rt.json').read().strip().splitlines()[-1])
print('rt$rt', d['value'], d['ms_per_step'], d['gpu_launches'], 'head_fwd', d['kernels']['head_fwd'], 'head_bwd', d['kernels']['head_bwd'])
PY
done
ADVMIL_RLIP_CHAIN=0 timeout 600 python bench.py --no-extra-legs --no-cpu-baseline > gpurun_out/bench_s14_off.json 2> gpurun_out/bench_s14.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_s14_off.json').read().strip().splitlines()[-1])
print('off', d['value'], d['ms_per_step'], d['gpu_launches'], 'head_fwd', d['kernels']['head_fwd'])
PY
