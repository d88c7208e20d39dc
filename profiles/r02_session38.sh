cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu_vl.json 2> gpurun_out/bench_8gpu_vl.err; echo "bench8 rc=$?"
tail -2 gpurun_out/bench_8gpu_vl.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_8gpu_vl.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('h2d_gbs_per_gpu'), d['e2e'].get('h2d_bytes_per_step'))
print({k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in (d.get('configs') or {}).items()}, d.get('sustained',{}).get('value'))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
