cd $GRAFT_REPO_ROOT
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke | cut -c1-200
timeout 200 python bench.py --steps 20 --warmup 5 --no-extra-legs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],4), d['gpu_launches'], d['clocks'])"
