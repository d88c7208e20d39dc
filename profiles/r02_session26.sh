cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -8 | cut -c1-300
timeout 300 python profiles/esat_bench.py --modes bf16 --steps 10 2>&1 | grep "^{" | cut -c1-200
