cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_esat.py tests/test_gpu_attention.py -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -8 | cut -c1-300
timeout 200 python -m pytest tests/test_gpu_step.py tests/test_gpu_dropin_loop.py -q -m gpu -k "esat" 2>&1 | grep -v "^  \|Warning\|^$" | tail -5 | cut -c1-300
for ov in 1 0; do
ADVMIL_ESAT_OVERLAP=$ov timeout 200 python profiles/esat_bench.py --modes bf16 --steps 10 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('overlap=$ov', {k:d[k] for k in d if k in ('ms_per_call','ms_per_step','bags_per_s')})"
done
