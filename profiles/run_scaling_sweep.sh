#!/bin/bash
# Scaling sweep of BASELINE configs[2] / [4]: global batch 512..8192 over G GPUs of one box (per-GPU batch = global / G),
# one bench.py line per point.  Usage: profiles/run_scaling_sweep.sh G "512 1024 ..." out.jsonl
G=$1; BATCHES=$2; OUT=$3
for gb in $BATCHES; do
  b=$((gb / G))
  if [ "$G" = "1" ]; then
    python bench.py --gpus 1 --batch $b --steps 5 --warmup 3 --no-cpu-baseline --no-stages --no-extra >> $OUT 2>> $OUT.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
      bench.py --gpus $G --batch $b --steps 5 --warmup 3 --no-cpu-baseline --no-stages --no-extra >> $OUT 2>> $OUT.err
  fi
done
