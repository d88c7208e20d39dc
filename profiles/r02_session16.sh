cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:rlip_chain' -s 6 -c 2 -f -o gpurun_out/r02_chain python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > /dev/null 2> gpurun_out/ncu.err; echo "ncu rc=$?"
tail -3 gpurun_out/ncu.err
