#!/bin/bash
# ncu captures of the stream kernel in population mode: a forward stage and the heaviest Adam stage
set -u
mkdir -p gpurun_out
CMD="python tools/throughput_stages.py ${AGENTS:-128} ${BATCH:-256}"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --cache-control none --import-source on -k regex:sac_stream_kernel -s ${SKIP_FWD:-22} -c 1 -f -o gpurun_out/r02_stream_fwd $CMD > gpurun_out/ncu_fwd.log 2>&1
echo "fwd rc=$?"
ncu --set full --clock-control none --cache-control none --import-source on -k regex:sac_stream_kernel -s ${SKIP_ADAM:-30} -c 1 -f -o gpurun_out/r02_stream_adam $CMD > gpurun_out/ncu_adam.log 2>&1
echo "adam rc=$?"
tail -3 gpurun_out/ncu_fwd.log gpurun_out/ncu_adam.log
