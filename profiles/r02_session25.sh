cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py -q -m gpu -k "transport or feeder" 2>&1 | grep -v "^  \|Warning\|^$" | tail -8 | cut -c1-300
for tr in vl; do
timeout 600 python bench.py --no-extra-legs --no-cpu-baseline --transport $tr > gpurun_out/bench_s25_$tr.json 2> gpurun_out/bench_s25.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_s25_$tr.json').read().strip().splitlines()[-1])
print('$tr', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['e2e']['h2d_gbs_per_gpu'])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:decode' -c 6 --csv --log-file gpurun_out/vl_decode.csv python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > /dev/null 2>&1
grep decode gpurun_out/vl_decode.csv | awk -F'","' '{print $5, $NF}' | cut -c1-80
