cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_step.py -q -m gpu -k "c_fused_esat" -x 2>&1 | grep -v "^  \|Warning\|^$" | tail -30 | cut -c1-400
