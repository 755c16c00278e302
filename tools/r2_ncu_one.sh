#!/bin/bash
# one ncu --set full capture of the stream kernel: SKIP = launch index, OUT = report name
set -u
mkdir -p gpurun_out
CMD="python tools/throughput_stages.py ${AGENTS:-128} ${BATCH:-256}"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --cache-control none --import-source on -k regex:${KERNEL:-sac_stream_kernel} -s ${SKIP:-29} -c ${COUNT:-1} -f -o gpurun_out/${OUT:-r02_stream_adam} $CMD > gpurun_out/ncu_one.log 2>&1
echo "ncu rc=$?"
