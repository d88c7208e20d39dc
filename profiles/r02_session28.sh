cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_esat.py tests/test_gpu_step.py -q -m gpu -k "esat" 2>&1 | grep -v "^  \|Warning\|^$" | tail -6 | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:ln_relu_mean16' -c 10 --csv --log-file gpurun_out/esat_ln_list.csv python profiles/esat_bench.py --modes bf16 --steps 3 > /dev/null 2>&1
grep mean16 gpurun_out/esat_ln_list.csv | awk -F'","' '{print substr($5,1,40), $NF}'
timeout 300 python profiles/esat_bench.py --modes bf16 --steps 10 2>&1 | grep "^{" | cut -c1-160
