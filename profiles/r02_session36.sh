cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_attention.py -q -m gpu -x 2>&1 | grep -v "^  \|Warning\|^$" | tail -25 | cut -c1-300
timeout 200 python -m pytest tests/test_gpu_esat.py -q -m gpu -x 2>&1 | grep -v "^  \|Warning\|^$" | tail -25 | cut -c1-300
