mkdir -p gpurun_out
G=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 300 --warmup 20 --no-cpu > gpurun_out/bench_g$G.log 2> gpurun_out/bench_g$G.err; echo "bench$G rc=$?"
grep "rank" gpurun_out/bench_g$G.err | tr '\n' ' '; echo
python - <<P
import json
for l in open("gpurun_out/bench_g$G.log"):
    if l.startswith("{"):
        d=json.loads(l); print("n_gpus", d["n_gpus"], "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
P
