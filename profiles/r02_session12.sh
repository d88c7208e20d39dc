cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v "^  \|Warning\|^$" | tail -30 | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_s12.json 2> gpurun_out/bench_s12.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_s12.err
