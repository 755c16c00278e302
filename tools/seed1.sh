timeout 900 python -X faulthandler -m pytest tests/test_gpu_replay.py -m gpu -q -x -p no:cacheprovider --timeout=300 --durations=5 2>&1 | tail -12
RANK=1 WORLD_SIZE=1 python - <<'P'
import os, sys, ctypes
sys.path.insert(0, os.getcwd())
import bench, numpy as np
import humanoid_walking_with_sac_b200 as hw
N = hw._native; lib = N.lib()
for seed in (1,):
    agent = bench.build_agent(hw, 0, "staged", "bf16x3", seed=seed)
    h = agent._h
    for _ in range(20): N.check(lib.sacb_per_step(h, 256, None, N.NO_LOSS_READBACK))
    agent.synchronize()
    ms = ctypes.c_float()
    N.check(lib.sacb_timer_start(h))
    for _ in range(200): N.check(lib.sacb_per_step(h, 256, None, N.NO_LOSS_READBACK))
    N.check(lib.sacb_timer_stop(h, ctypes.byref(ms)))
    st = N.PerStats(); N.check(lib.sacb_per_get_stats(h, 0, ctypes.byref(st)))
    print("seed", seed, "ms/step", ms.value / 200, "n_fine", st.n_fine, "flagged", st.n_flagged, "exact", st.n_exact_fallbacks, "total", st.total_f32)
P
