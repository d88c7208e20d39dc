cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for m in 0 4 0 4; do
  ADVMIL_RLIP_CHAIN_MMA=$m timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0.3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('mma=$m', round(d['value']), {k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in d['configs'].items()}, {k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in d['modes'].items()})"
done
