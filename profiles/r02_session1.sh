set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r02a.err
ADVMIL_TC_CLUSTER=2 timeout 300 python bench.py --no-extra-legs --no-cpu-baseline --steps 20 > gpurun_out/bench_cl2.json 2>/dev/null; echo "cl2 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 600 python -m pytest tests/test_gpu_handler.py -q -m gpu 2>&1 | tail -15
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:tc_rows_kernel<__nv_bfloat16' -s 15 -c 5 -f -o gpurun_out/r02_rows_bf16 python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > /dev/null 2> gpurun_out/ncu.err; echo "ncu rc=$?"
tail -5 gpurun_out/ncu.err
ls -la gpurun_out
