cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:mha_' -c 12 --csv --log-file gpurun_out/esat_attn_list.csv python profiles/esat_bench.py --modes bf16 --steps 3 > /dev/null 2>&1
grep "mha" gpurun_out/esat_attn_list.csv | awk -F'","' '{print substr($5,1,60), $NF}'
timeout 200 python profiles/esat_bench.py --modes bf16 --steps 10 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in d if k in ('what','ms_per_call','ms_per_step','bags_per_s')})"
