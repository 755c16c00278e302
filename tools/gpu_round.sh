#!/bin/bash
# One GPU round trip: build check -> GEMM self test -> parity tests -> stage trace -> bench.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
timeout 300 python -X faulthandler -m pytest tests/test_gpu_gemm.py -m gpu -q -x -p no:cacheprovider --timeout=120 > gpurun_out/gemm.log 2>&1; rc=$?; echo "gemm rc=$rc" | tee -a gpurun_out/summary.log
tail -5 gpurun_out/gemm.log
if [ $rc -ne 0 ] && [ "${FORCE:-0}" != "1" ]; then exit 0; fi
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --timeout=300 --deselect tests/test_gpu_gemm.py > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.log
tail -30 gpurun_out/pytest.log
if [ "${TRACE:-1}" = "1" ]; then
SACB_TIMELINE=1 timeout 300 python tools/trace_stages.py > gpurun_out/trace.log 2>&1; echo "trace rc=$?" | tee -a gpurun_out/summary.log
grep "timeline\|\[trace\] stage" gpurun_out/trace.log | cut -c1-230
fi
timeout 900 python bench.py --steps 300 --warmup 20 ${BENCH_ARGS:-} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.log
tail -5 gpurun_out/bench.err
python - <<P
import json
for l in open("gpurun_out/bench.log"):
    if l.startswith("{"):
        d=json.loads(l); print("ms/step", d["ms_per_step"], "value", d["value"], "upd", d["roofline"]["ms_per_launch_sum"], "e2e", d["e2e"]["value"], "per", d["roofline"]["per_sample"]["ms_per_call"], "launches", d["gpu_launches"], "cpu", d["cpu_baseline"])
P
