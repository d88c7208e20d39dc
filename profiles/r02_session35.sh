cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_step.py -q -m gpu -k "packed_file or transport or feeder" 2>&1 | grep -v "^  \|Warning\|^$" | tail -6 | cut -c1-250
