#!/bin/bash
set -u
timeout 900 python -X faulthandler -m pytest tests/test_gpu_sharded.py tests/test_gpu_update_parity.py -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=300 2>&1 | tail -2
timeout 300 python tools/throughput_stages.py 128 256 2>&1 | grep "THROUGHPUT\|stage_us"
for ph in 0 1; do SACB_TIME_DP_PHASE=$ph timeout 300 python tools/throughput_stages.py 1 8192 2>&1 | grep "stage_us"; done
timeout 300 python tools/dp_bench.py 8192 30 2>&1 | grep DP_BENCH | cut -c1-160
timeout 300 python tools/dp_bench.py 65536 10 2>&1 | grep DP_BENCH | cut -c1-160
timeout 300 python tools/population_bench.py 2>&1 | tail -3 | cut -c1-300
