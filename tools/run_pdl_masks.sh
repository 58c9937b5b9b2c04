# sweep of MVAE_PDL bit masks (bit per launch: 0 enc_fwd, 1 tail_fwd, 2 dec_fwd, 3 dec_bwd, 4 tail_bwd, 5 enc_bwd)
for v in ${MASKS:-0 1 2 4 8 16 32 0}; do
  MVAE_PDL=$v timeout 100 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MVAE_PDL=$v', 'ms_per_step', round(d['ms_per_step'] * 1e3, 2), 'us  e2e', round(d['e2e']['value']))"
done
