cd $GRAFT_REPO_ROOT
timeout 150 python -m pytest tests/test_gpu_attention.py -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -25 | cut -c1-300
