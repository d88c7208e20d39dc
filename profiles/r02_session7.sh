cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r02_bf16_parity.json
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -30 | cut -c1-300
timeout 600 python -m pytest tests -q -m gpu -p no:randomly tests/test_gpu_tf32x3.py tests/test_gpu_step.py 2>&1 | tail -3
