#!/bin/bash
# 2+ GPU check of the data-parallel mode: parity (G ranks x B/G == 1 GPU x B) and throughput, fused peer exchange vs NCCL
set -u
G=${G:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/dp_check.py > gpurun_out/dp_check_$G.log 2>&1; echo "dp_check rc=$?"; grep "DP_CHECK" gpurun_out/dp_check_$G.log | cut -c1-400; tail -3 gpurun_out/dp_check_$G.log | cut -c1-300
for GB in 8192 65536; do
  for X in peer nccl; do
    SACB_DP_EXCHANGE=$X timeout 300 $TR tools/dp_bench.py $GB ${STEPS:-20} > gpurun_out/dp_bench_${G}_${GB}_$X.log 2>&1; echo "dp_bench $GB $X rc=$?"
    grep "DP_BENCH" gpurun_out/dp_bench_${G}_${GB}_$X.log | cut -c1-220
  done
done
