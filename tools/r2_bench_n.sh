#!/bin/bash
# bench.py under torchrun at G GPUs (the driver's launch line)
set -u
G=${G:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
if [ "$G" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps ${STEPS:-100} --warmup 5 ${BENCH_ARGS:-} > gpurun_out/bench_g$G.log 2> gpurun_out/bench_g$G.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $G --steps ${STEPS:-100} --warmup 5 ${BENCH_ARGS:-} > gpurun_out/bench_g$G.log 2> gpurun_out/bench_g$G.err
fi
echo "bench rc=$?"
grep -v "^\[rank\|^W1\|^\*\*\*" gpurun_out/bench_g$G.err | tail -8 | cut -c1-300
python - <<P
import json
for l in open("gpurun_out/bench_g$G.log"):
    if l.startswith("{"):
        d=json.loads(l); print("N", d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "k8", round(d["e2e"]["batched_k8"]["value"]), "launches", d["gpu_launches"], "clocks", d["clocks"])
        s=d["sharded"]
        if s and "population" in s:
            p=s["population"]; print("POP", p["agents_total"], "agents", round(p["agent_updates_per_s"]), "upd/s", round(p["algorithmic_TFLOPs_per_gpu"],1), "TF/gpu", "hbm frac", round(p["roofline"]["frac"],3))
            for x in s["data_parallel"]: print("DP", x["global_batch"], "ms", round(x["ms_per_update"],3), "upd/s", round(x["updates_per_s"],1), "TF/gpu", round(x["algorithmic_TFLOPs_per_gpu"],1), "frac", round(x["roofline_frac_per_gpu"],3), "exchange share", round(x["exchange_share"],3))
        else: print("sharded", s)
P
