cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 1200 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest -q -m gpu "tests/test_gpu_tf32x3.py::test_x3_fused_step_vs_reference_golden" -x > gpurun_out/memcheck_x3.log 2>&1
grep -v "^$" gpurun_out/memcheck_x3.log | grep -B2 -A28 "Invalid" | head -120
tail -5 gpurun_out/memcheck_x3.log
