cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_library.py tests/test_gpu_parity_modules.py tests/test_gpu_step.py tests/test_gpu_dropin_loop.py tests/test_gpu_handler.py -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -30 | cut -c1-300
python profiles/prof_dropin.py fp32 2>&1 | grep "^{"
python profiles/prof_dropin.py bf16 2>&1 | grep "^{"
