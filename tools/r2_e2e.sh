#!/bin/bash
# replay / agent tests + a short bench line (no CPU arm, no sharded block): e2e and value after a replay-path change
set -u
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest ${TESTS:-tests/test_gpu_replay.py tests/test_gpu_agent_api.py} -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
grep -v "^\.\|^$\|^s" gpurun_out/pytest.log | tail -15 | cut -c1-300
timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu --no-sharded > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open("gpurun_out/bench_quick.log").read().strip().splitlines()[-1])
print("value", round(l["value"],1), "ms/step", round(l["ms_per_step"],4), "e2e", round(l["e2e"]["value"],1), "k8", round(l["e2e"].get("batched_k8",{}).get("value",0),1), "per_sample ms", l["roofline"].get("per_sample",{}).get("ms_per_call"))
PY
