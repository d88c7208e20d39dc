cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -14
nproc; lscpu | grep -i "numa node" | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench8 rc=$?"
tail -3 gpurun_out/bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_8gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('h2d_gbs_per_gpu'), d['e2e'].get('host_numa_binding'))
print(d.get('configs'), d.get('sustained'))
PY
timeout 300 python -m pytest tests/test_gpu_dp.py -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r02_test_gpu_dp.log
