#!/bin/bash
# usage: tools/run_scaling.sh N OUTDIR  -- runs, on N GPUs of one box: the headline bench (scene sharding, weak scaling),
# config c3 (recurrent fusion over x-slabs) and config c5 (32 scenes, 96^3, reduce-scatter by slab / slab owner; strong scaling)
N=$1; OUT=$2; mkdir -p $OUT
if [ "$N" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"; fi
$L bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-convlstm > $OUT/scene_$N.json 2> $OUT/scene_$N.err
$L bench.py --gpus $N --strategy lstm_slab --scenes 1 --steps 2 --warmup 1 > $OUT/c3_lstm_slab_$N.json 2> $OUT/c3_lstm_slab_$N.err
for st in view_reduce_scatter slab_owner view_allreduce; do
  $L bench.py --gpus $N --strategy $st --scenes 32 --nvox 96 --steps 3 --warmup 3 > $OUT/c5_${st}_$N.json 2> $OUT/c5_${st}_$N.err
done
tail -c 300 $OUT/*_$N.err
for f in $OUT/*_$N.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','scaling','useful_tflops')})
except Exception as e: print('ERR',e)"; done
