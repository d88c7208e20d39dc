cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-extra-legs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms_per_launch']*1000,1) for k,v in d['kernels'].items() if k in ('ln_bwd','head_fwd','head_bwd')})"; }
run A=0
run ADVMIL_LN_BWD_OCC=5
run ADVMIL_RLIP_CHAIN_STAGGER=5000
run ADVMIL_RLIP_CHAIN_STAGGER=10000
run A=0
run ADVMIL_LN_BWD_OCC=5
for s in 0 5000 10000; do
ADVMIL_RLIP_CHAIN_STAGGER=$s timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:rlip_chain|ln_pool_bwd128' -s 9 -c 3 --csv --log-file gpurun_out/knob.csv python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > /dev/null 2>&1
echo "stagger $s: $(tail -3 gpurun_out/knob.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done
ADVMIL_LN_BWD_OCC=5 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:ln_pool_bwd128' -s 3 -c 2 --csv --log-file gpurun_out/knob.csv python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > /dev/null 2>&1
echo "ln occ5: $(tail -2 gpurun_out/knob.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
