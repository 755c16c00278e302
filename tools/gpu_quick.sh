#!/bin/bash
# quick GPU check: replay + agent tests, then the bench both ways
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_replay.py tests/test_gpu_agent_api.py -m gpu -q -x -p no:cacheprovider --timeout=300 > gpurun_out/quick_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/quick_pytest.log
for f in "" "--no-pipeline"; do
  timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu $f > gpurun_out/quick_bench$f.log 2> gpurun_out/quick_bench$f.err; echo "bench $f rc=$?"
  tail -3 gpurun_out/quick_bench$f.err
  python - <<P
import json
for l in open("gpurun_out/quick_bench$f.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$f", "ms/step", d["ms_per_step"], "upd", d["roofline"]["ms_per_launch_sum"], "e2e", d["e2e"]["value"], "per", d["roofline"]["per_sample"]["ms_per_call"], "launches", d["gpu_launches"])
P
done
