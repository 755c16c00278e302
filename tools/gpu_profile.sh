#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the bench command, (2) the stage kernels of one update with DRAM / L2 traffic,
# (3) full captures of a forward stage and of the heaviest Adam stage, (4) the PER sampler kernels.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 12 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -c 520 \
    --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
PCMD="python tools/profile_update.py bf16x3 staged 3"
$PCMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --cache-control none -k regex:sac_update_kernel -s 52 -c 26 --csv --log-file gpurun_out/update_stages.csv $PCMD > gpurun_out/ncu3.log 2>&1
echo "stages rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sac_update_kernel -s 54 -c 1 -f -o gpurun_out/prof_fwd $PCMD > gpurun_out/ncu1.log 2>&1
echo "fwd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sac_update_kernel -s 64 -c 1 -f -o gpurun_out/prof_adam $PCMD > gpurun_out/ncu2.log 2>&1
echo "adam rc=$?"
ncu --set full --clock-control none --cache-control none --import-source on -k regex:per_ -s 30 -c 3 -f -o gpurun_out/per_full python tools/per_profile.py 1000000 10 0 > gpurun_out/per_ncu2.log 2>&1
echo "per rc=$?"
