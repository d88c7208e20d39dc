cd $GRAFT_REPO_ROOT
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 1400 --csv --log-file gpurun_out/launches_r02_esat.csv python bench.py --backbone patch --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "list rc=$?"; wc -l gpurun_out/launches_r02_esat.csv
