mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu > gpurun_out/bench_g2.log 2> gpurun_out/bench_g2.err; echo "bench2 rc=$?"
grep "rank" gpurun_out/bench_g2.err
# same two ranks, but each process sees only its own GPU
cat > /tmp/wrap.sh <<'W'
#!/bin/bash
export CUDA_VISIBLE_DEVICES=$LOCAL_RANK
export LOCAL_RANK=0
exec python "$@"
W
chmod +x /tmp/wrap.sh
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 --no-python /tmp/wrap.sh bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu > gpurun_out/bench_g2w.log 2> gpurun_out/bench_g2w.err; echo "wrapped rc=$?"
grep "rank" gpurun_out/bench_g2w.err; tail -2 gpurun_out/bench_g2w.err | cut -c1-300
