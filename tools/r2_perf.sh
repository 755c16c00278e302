#!/bin/bash
# throughput-mode stage profiles + single-agent update time
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
python tools/ab_update.py 2>&1 | grep AB_UPDATE
${TRACE:+env SACB_TRACE=1} timeout 300 python tools/throughput_stages.py 1 8192 > gpurun_out/stages_b8192.log 2>&1; echo "stages8192 rc=$?"
grep "THROUGHPUT\|stage_us" gpurun_out/stages_b8192.log
${TRACE:+env SACB_TRACE=1} timeout 300 python tools/throughput_stages.py 128 256 > gpurun_out/stages_pop128.log 2>&1; echo "stagespop rc=$?"
grep "THROUGHPUT\|stage_us" gpurun_out/stages_pop128.log
if [ "${NOSTREAM:-0}" = "1" ]; then
SACB_NO_STREAM=1 timeout 300 python tools/throughput_stages.py 1 8192 2>&1 | grep "THROUGHPUT\|stage_us"
SACB_NO_STREAM=1 timeout 300 python tools/throughput_stages.py 128 256 2>&1 | grep "THROUGHPUT\|stage_us"
fi
if [ "${DP:-1}" = "1" ]; then
timeout 300 python tools/dp_bench.py 8192 20 2>&1 | grep DP_BENCH
timeout 300 python tools/dp_bench.py 65536 6 2>&1 | grep DP_BENCH
fi
