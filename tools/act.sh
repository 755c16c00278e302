python tools/act_latency.py
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_agent_api.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
