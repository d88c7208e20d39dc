cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:mha_' -c 24 --csv --log-file gpurun_out/esat_attn_list.csv python profiles/esat_bench.py --modes bf16 --steps 3 > /dev/null 2>&1
grep mha gpurun_out/esat_attn_list.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
