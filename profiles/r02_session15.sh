cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 1500 --csv --log-file gpurun_out/launches_r02c.csv python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > gpurun_out/b.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/b.log
wc -l gpurun_out/launches_r02c.csv
