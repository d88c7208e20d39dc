cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python profiles/sanitize_r02_new_kernels.py > gpurun_out/memcheck_r02_new.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/memcheck_r02_new.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -s 700 -c 260 --csv --log-file gpurun_out/launches_r02_esat.csv python profiles/esat_bench.py --modes bf16 --steps 4 > /dev/null 2>&1; echo "list rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:mha_' -s 9 -c 3 -f -o gpurun_out/r02_attention python profiles/esat_bench.py --modes bf16 --steps 3 > /dev/null 2>&1; echo "ncu rc=$?"
