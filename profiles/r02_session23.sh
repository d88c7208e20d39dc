cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python profiles/esat_bench.py --modes bf16 --steps 10 2>&1 | grep "^{" | tee gpurun_out/r02_esat_generator_bench.jsonl
timeout 600 python bench.py --backbone patch --no-cpu-baseline > gpurun_out/bench_r02_esat.json 2> gpurun_out/bench_r02_esat.err; echo "esat bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_esat.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('gpu_launches'), json.dumps(d.get('kernels'))[:1500])
PY
