cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_esat.py -q -m gpu -x 2>&1 | grep -v "^  \|Warning\|^$" | tail -25 | cut -c1-300
timeout 300 python profiles/esat_bench.py --modes bf16 --steps 10 2>&1 | tail -5
