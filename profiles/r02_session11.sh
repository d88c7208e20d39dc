cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -20
nproc; lscpu | grep -i "numa\|socket\|model name" | head
timeout 600 python -m pytest tests/test_gpu_dp.py tests/test_gpu_library.py tests/test_gpu_parity_modules.py -q -m gpu 2>&1 | tail -12 | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"
tail -3 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('h2d_gbs_per_gpu'), d['e2e'].get('host_numa_binding'))
print(d.get('configs'), d.get('sustained'))
PY
