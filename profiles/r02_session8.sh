cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest -q -m gpu "tests/test_gpu_tf32x3.py::test_x3_fused_step_vs_reference_golden" 2>&1 | tail -3
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest -q -m gpu "tests/test_gpu_tf32x3.py::test_x3_fused_step_vs_reference_golden" -x 2>&1 | grep -v "^$" | grep -A25 "Invalid\|=========.*ERROR\|Error" | head -80
