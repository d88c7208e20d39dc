cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_parity_modules.py tests/test_gpu_step.py tests/test_gpu_tf32x3.py -q -m gpu -x 2>&1 | grep -v "^  \|Warning\|^$" | tail -8 | cut -c1-300
echo "tests $(( $(date +%s) - t0 )) s"
for m in 1 0; do
  ADVMIL_RLIP_CHAIN_MMA=$m timeout 300 python bench.py --steps 30 --warmup 5 --no-extra-legs --no-cpu-baseline > gpurun_out/bench_s50_mma$m.json 2> gpurun_out/bench_s50_mma$m.err; echo "bench mma=$m rc=$?"
done
python - <<'PY'
import json
for m in (1, 0):
    try:
        d = json.loads(open(f'gpurun_out/bench_s50_mma{m}.json').read().strip().splitlines()[-1])
        print('mma', m, d['value'], d['ms_per_step'], 'launches', d['gpu_launches'])
        prof = d.get('kernels') or {}
        print({k: v for k, v in prof.items() if 'head' in k})
    except Exception as ex:
        print('mma', m, 'failed', ex)
PY
echo "total $(( $(date +%s) - t0 )) s"
