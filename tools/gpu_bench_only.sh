#!/bin/bash
# timeline of one update + the bench line (no tests)
mkdir -p gpurun_out
SACB_TIMELINE=1 timeout 300 python tools/trace_stages.py > gpurun_out/trace.log 2>&1
grep "timeline" gpurun_out/trace.log | awk '{print $2, $6, $NF}' | tr '\n' ';'; echo
timeout 900 python bench.py --steps 300 --warmup 20 ${BENCH_ARGS:-} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.err
python - <<P
import json
for l in open("gpurun_out/bench.log"):
    if l.startswith("{"):
        d=json.loads(l); print("ms/step", d["ms_per_step"], "value", d["value"], "upd", d["roofline"]["ms_per_launch_sum"], "e2e", d["e2e"]["value"], "per", d["roofline"]["per_sample"]["ms_per_call"], "launches", d["gpu_launches"], "cpu", d["cpu_baseline"])
P
