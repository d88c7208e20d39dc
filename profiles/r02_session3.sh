cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --deselect "tests/test_gpu_handler.py::test_unmodified_handler_epoch_in_the_fast_modes[tf32x3]" 2>&1 | grep -v "^  \|Warning\|^$" | tail -60
