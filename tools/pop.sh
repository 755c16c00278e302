mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests/test_gpu_gemm.py -m gpu -q -x -p no:cacheprovider --timeout=120 2>&1 | tail -4
for a in 8 64; do timeout 300 python tools/population_bench.py $a 10 2>&1 | tail -1 | cut -c1-140; done
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_update_parity.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
BENCH_ARGS=--no-cpu bash tools/gpu_bench_only.sh
