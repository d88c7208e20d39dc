cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dp.py -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -3 | cut -c1-300 | tee gpurun_out/r02_test_gpu_dp_final.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02_2gpu_final.json 2> gpurun_out/bench_r02_2gpu_final.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_2gpu_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], {k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in d.get('configs',{}).items()})
PY
