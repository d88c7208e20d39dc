cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -s 300 -c 170 --csv --log-file gpurun_out/launches_r02_esat.csv python profiles/esat_bench.py --modes bf16 --steps 4 > /dev/null 2>&1; echo rc=$?
