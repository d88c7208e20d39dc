cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 4 5 6; do
  ADVMIL_RLIP_CHAIN_MMA=$v timeout 300 python -m pytest tests/test_gpu_parity_modules.py tests/test_gpu_step.py tests/test_gpu_tf32x3.py -q -m gpu -x 2>&1 | grep -v "^  \|Warning\|^$" | tail -2 | cut -c1-200
  ADVMIL_RLIP_CHAIN_MMA=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:rlip_chain' -s 6 -c 4 --csv --log-file gpurun_out/chain_v$v.csv python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > /dev/null 2>&1
  echo "variant $v: $(tail -4 gpurun_out/chain_v$v.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
  ADVMIL_RLIP_CHAIN_MMA=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-extra-legs --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['kernels']['head_fwd']['ms_per_launch'])"
done
