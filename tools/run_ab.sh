# A/B of an environment switch inside one box: bench.py twice per setting, interleaved
for rep in 1 2; do for v in 0 1; do
  env $1=$v timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1=$v', 'ms_per_step', round(d['ms_per_step'] * 1e3, 2), 'us  e2e', round(d['e2e']['value']))"
done; done
