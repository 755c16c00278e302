#!/bin/bash
# One GPU round trip: self test -> parity tests -> bench (+ optional ncu).  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
timeout 300 python -X faulthandler -m pytest tests/test_gpu_gemm.py -m gpu -q -p no:cacheprovider --timeout=120 > gpurun_out/gemm.log 2>&1; echo "gemm rc=$?" | tee -a gpurun_out/summary.log
tail -5 gpurun_out/gemm.log
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --timeout=300 --deselect tests/test_gpu_gemm.py > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.log
tail -30 gpurun_out/pytest.log
if [ -f bench.py ]; then
  timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.log
  tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
fi
if [ "${PROFILE:-0}" = "1" ]; then
  rm -f gpurun_out/*.ncu-rep
  python tools/profile_update.py tf32x3 staged 3 > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sac_update_kernel -s 26 -c 1 -o gpurun_out/prof_fwd python tools/profile_update.py tf32x3 staged 3 > gpurun_out/ncu1.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:sac_update_kernel -s 35 -c 1 -o gpurun_out/prof_bwd python tools/profile_update.py tf32x3 staged 3 > gpurun_out/ncu2.log 2>&1
fi
