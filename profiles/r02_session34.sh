cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -8 | cut -c1-300
