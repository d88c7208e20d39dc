cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^  \|Warning\|^$" | tail -4 | cut -c1-300
for w in 1 0 1 0; do
  ADVMIL_PE_WIDE=$w timeout 300 python bench.py --steps 30 --warmup 5 --no-extra-legs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('wide=$w', round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms_per_launch']*1000,1) for k,v in d['kernels'].items() if k in ('proj_embed_fwd','head_bwd','head_fwd','gate_fwd')})"
done
