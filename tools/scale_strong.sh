#!/bin/bash
# Strong-scaling curve (BASELINE config 5 shape: a FIXED set of Drugs-shaped molecules x 2 samples, LPT-sharded): run under
# gpurun --gpus N.  usage: tools/scale_strong.sh <N> <total molecules> <out.json>
N=$1; M=$2; OUT=$3
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 1 --warmup 1 --total-mols $M --no-cpu-baseline > $OUT
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 1 --warmup 1 --total-mols $M --no-cpu-baseline | grep '^{' > $OUT
fi
python -c "import json,sys; d=json.load(open('$OUT')); print('N', d['n_gpus'], 'conformers/s', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'imbalance', d['shard_imbalance'], 'ms/step', round(d['ms_per_step']))"
