cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -15 | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_r02_default.json 2> gpurun_out/bench_r02_default.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 1500 --csv --log-file gpurun_out/launches_r02_bf16.csv python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > gpurun_out/b.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:tc_rows_kernel<__nv_bfloat16|tc_wgrad_kernel<__nv_bfloat16|pool_gate_bwd_kernel<__nv_bfloat16|ln_pool_bwd128|seg_pool_partial_kernel<__nv_bfloat16|rlip_chain' -s 30 -c 16 -f -o gpurun_out/r02_hot python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > /dev/null 2> gpurun_out/ncu.err; echo "ncu full rc=$?"
