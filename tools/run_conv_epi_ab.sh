for v in 0 1 0 1; do for w in celeba multimnist; do
MVAE_CONV_EPI_STATS=$v timeout 300 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w epi_stats=$v', 'ms/step', round(d['ms_per_step'], 4), 'value', round(d['value']), 'launches', d.get('gpu_launches_per_step'))"
done; done
