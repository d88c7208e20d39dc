cd $GRAFT_REPO_ROOT
for ov in 0 1; do
ADVMIL_ESAT_OVERLAP=$ov timeout 300 python profiles/esat_bench.py --modes bf16 --steps 10 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('overlap=$ov', {k:d[k] for k in d if k in ('what','ms_per_call','ms_per_step')})" | grep -v Module
done
