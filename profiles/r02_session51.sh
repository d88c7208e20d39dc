cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:rlip_chain' -s 6 -c 4 -f -o gpurun_out/r02_chain_mma python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > /dev/null 2>&1; echo "ncu rc=$?"
ADVMIL_RLIP_CHAIN_MMA=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:rlip_chain' -s 6 -c 4 --csv --log-file gpurun_out/chain_ffma_list.csv python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > /dev/null 2>&1; echo "list rc=$?"
tail -4 gpurun_out/chain_ffma_list.csv | cut -c1-60,200-
