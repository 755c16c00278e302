#!/bin/bash
# ncu --set full of one forward stage of the stream kernel at B = 65536 (policy on 2B rows + both critics, K = N = 512): 128 x 128 against 128 x 256 tiles
set -u
mkdir -p gpurun_out
for v in 0 1; do
  CMD="python tools/throughput_stages.py 1 ${BATCH:-65536}"
  SACB_STREAM_N256_MIN=$v $CMD > gpurun_out/ncu_plain.log 2>&1 &&
  SACB_STREAM_N256_MIN=$v ncu --set full --clock-control none --cache-control none --import-source on -k regex:sac_stream_kernel -s 1 -c 1 -f -o gpurun_out/r02_stream_fwd_b65536_n256_$v $CMD > gpurun_out/ncu_wide_$v.log 2>&1
  echo "ncu rc=$? (N256_MIN=$v)"
done
