#!/bin/bash
# data-parallel step at G GPUs with the 128 x 256 stream tile off (0) / on from 1, 2 waves
set -u
G=${G:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
for v in 0 1 2; do
  for GB in 8192 65536; do
    SACB_STREAM_N256_MIN=$v timeout 300 $TR tools/dp_bench.py $GB 30 2>&1 | grep "DP_BENCH" | cut -c1-150 | sed "s/^/N256_MIN=$v /"
  done
done
