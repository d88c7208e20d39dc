cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r02_bf16_parity.json
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -40 | cut -c1-300
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:tc_rows_kernel<__nv_bfloat16' -s 15 -c 5 -f -o gpurun_out/r02b_rows_bf16 python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > /dev/null 2> gpurun_out/ncu.err; echo "ncu rc=$?"
