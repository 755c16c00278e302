mkdir -p gpurun_out
G=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 200 --warmup 20 > gpurun_out/bench_g$G.log 2> gpurun_out/bench_g$G.err; echo "bench$G rc=$?"; tail -2 gpurun_out/bench_g$G.err
python - <<P
import json
for l in open("gpurun_out/bench_g$G.log"):
    if l.startswith("{"):
        d=json.loads(l); print("n_gpus", d["n_gpus"], "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
P
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $G --steps 5 --warmup 1 2>&1 | tail -1 | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 tools/dp_check.py 2>&1 | grep DP_CHECK | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29514 tools/dp_bench.py 8192 20 2>&1 | grep DP_BENCH
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29515 tools/dp_bench.py 65536 5 2>&1 | grep DP_BENCH
