cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:mha_fwd' -s 5 -c 2 -f -o gpurun_out/r02_attn_fwd python profiles/esat_bench.py --modes bf16 --steps 3 > /dev/null 2>&1; echo rc=$?
