cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for occ in 2 3; do
ADVMIL_RLIP_CHAIN_OCC=$occ timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:rlip_chain' -s 6 -c 4 --csv --log-file gpurun_out/chain_occ$occ.csv python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > /dev/null 2>&1
grep rlip gpurun_out/chain_occ$occ.csv | awk -F'","' '{print $NF}'
done
