cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests/test_gpu_rlip_chain.py -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -12 | cut -c1-300
