#!/bin/bash
# 128 x 256 stream tile: self tests, sharded parity tests, throughput A/B (SACB_STREAM_N256_MIN=0 = the 128 x 128 form)
set -u
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_gemm.py tests/test_gpu_sharded.py -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
grep -v "^\.\|^$\|^s" gpurun_out/pytest.log | tail -12 | cut -c1-300
for v in 0 3 ${EXTRA:-}; do
  echo "== SACB_STREAM_N256_MIN=$v"
  SACB_STREAM_N256_MIN=$v timeout 300 python tools/throughput_stages.py 1 8192 2>&1 | grep "THROUGHPUT\|stage_us"
  SACB_STREAM_N256_MIN=$v timeout 300 python tools/throughput_stages.py 1 65536 2>&1 | grep "THROUGHPUT\|stage_us"
  SACB_STREAM_N256_MIN=$v timeout 300 python tools/throughput_stages.py 128 256 2>&1 | grep "THROUGHPUT\|stage_us"
  SACB_STREAM_N256_MIN=$v timeout 300 python tools/dp_bench.py 8192 20 2>&1 | grep DP_BENCH | cut -c1-200
  SACB_STREAM_N256_MIN=$v timeout 300 python tools/dp_bench.py 65536 6 2>&1 | grep DP_BENCH | cut -c1-200
done
