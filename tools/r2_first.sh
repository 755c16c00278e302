#!/bin/bash
# round 2, first GPU round trip: build check, full GPU test suite, throughput-mode stage profiles, quick bench
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -X faulthandler -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider --timeout=300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest.log
SACB_TRACE=1 timeout 300 python tools/throughput_stages.py 1 8192 > gpurun_out/stages_b8192.log 2>&1; echo "stages8192 rc=$?"
grep "THROUGHPUT\|stage_us" gpurun_out/stages_b8192.log
SACB_TRACE=1 timeout 300 python tools/throughput_stages.py 128 256 > gpurun_out/stages_pop128.log 2>&1; echo "stagespop rc=$?"
grep "THROUGHPUT\|stage_us" gpurun_out/stages_pop128.log
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.err
python - <<P
import json
for l in open("gpurun_out/bench.log"):
    if l.startswith("{"):
        d=json.loads(l); print("ms/step", d["ms_per_step"], "value", d["value"], "upd", d["roofline"]["ms_per_launch_sum"], "e2e", d["e2e"]["value"], "k8", d["e2e"]["batched_k8"]["value"], "per", d["roofline"]["per_sample"]["ms_per_call"], "launches", d["gpu_launches"])
        print("cpu", d["cpu_baseline"]); print("eager", d["eager_cuda_baseline"]); print("sharded", json.dumps(d["sharded"])[:3000]); print("stage_us", d["roofline"]["stage_us"])
P
