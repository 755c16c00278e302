mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/dp_bench.py 8192 20 2>&1 | tail -2
timeout 300 python tools/dp_bench.py 65536 5 2>&1 | tail -2
