# A/B at 2 GPUs: column split of enc_bwd under data parallel (MVAE_CHAIN_SPLIT_DP) x grid of the decoder bucket's exchange kernel
i=0
IFS=","; for cfg in ${CFGS:-1 96,4 12,2 19}; do IFS=" "
set -- $cfg; i=$((i+1))
MVAE_CHAIN_SPLIT_DP=$1 MVAE_DP_DEC_BLOCKS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$i bench.py --gpus 2 --steps 40 --warmup 10 --no-extra 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dp_parts=$1 dec_blocks=$2', 'us/step', round(d['ms_per_step'] * 1e3, 2), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
