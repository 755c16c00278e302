#!/bin/bash
# single-GPU sweep of mid-size batches: stream-kernel tile width heuristic
for B in 1024 2048 4096 8192; do
  for WM in 0 3 100; do
    SACB_STREAM_WIDE_MIN=$WM python tools/throughput_stages.py 1 $B 2>&1 | grep THROUGHPUT | sed "s/^/wide_min=$WM /"
  done
done
for A in 16 32 64; do
  for WM in 0 3 100; do
    SACB_STREAM_WIDE_MIN=$WM python tools/throughput_stages.py $A 256 2>&1 | grep THROUGHPUT | sed "s/^/wide_min=$WM /"
  done
done
