cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_esat.py -q -m gpu -x 2>&1 | grep -v "^  \|Warning\|^$" | tail -8 | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:mha_fwd' -c 8 --csv --log-file gpurun_out/esat_attn_list.csv python profiles/esat_bench.py --modes bf16 --steps 3 > /dev/null 2>&1
grep mha gpurun_out/esat_attn_list.csv | awk -F'","' '{print $NF}' | tr '\n' ' '
