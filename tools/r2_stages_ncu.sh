#!/bin/bash
# Round-2 refresh of the per-stage evidence of ONE single-agent update (C2 shape, B = 256, staged launch): time, DRAM bytes, L2 bytes and
# tensor-pipe activity of the batch-staging kernel and each of the 25 stage kernels with warm caches, plus one full capture of a forward stage and of the heaviest
# Adam stage.  Output: gpurun_out/r2_update_stages.csv, r2_lat_fwd.ncu-rep, r2_lat_adam.ncu-rep
set -u
mkdir -p gpurun_out
PCMD="python tools/profile_update.py bf16x3 staged 3"
$PCMD > gpurun_out/r2_prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --cache-control none -k regex:sac_update_kernel -s 52 -c 26 --csv --log-file gpurun_out/r2_update_stages.csv $PCMD > gpurun_out/r2_ncu_stages.log 2>&1
echo "stages rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sac_update_kernel -s 54 -c 1 -f -o gpurun_out/r2_lat_fwd $PCMD > gpurun_out/r2_ncu_fwd.log 2>&1
echo "fwd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sac_update_kernel -s 64 -c 1 -f -o gpurun_out/r2_lat_adam $PCMD > gpurun_out/r2_ncu_adam.log 2>&1
echo "adam rc=$?"
grep -c sac_update_kernel gpurun_out/r2_update_stages.csv
