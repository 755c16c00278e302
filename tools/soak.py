"""Soak run of the pipelined learner step (C2, 1 M-row prioritized ring): step time per block of 2000 steps, PER statistics, losses."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import humanoid_walking_with_sac_b200 as hw
N = hw._native; lib = N.lib()
blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 10
agent = bench.build_agent(hw, 0, "staged", "bf16x3", seed=3)
h = agent._h
for _ in range(20):
    N.check(lib.sacb_per_step(h, 256, None, N.NO_LOSS_READBACK))
for blk in range(blocks):
    ms = ctypes.c_float()
    N.check(lib.sacb_timer_start(h))
    for _ in range(2000):
        N.check(lib.sacb_per_step(h, 256, None, N.NO_LOSS_READBACK))
    N.check(lib.sacb_timer_stop(h, ctypes.byref(ms)))
    st = N.PerStats(); N.check(lib.sacb_per_get_stats(h, 0, ctypes.byref(st)))
    losses = np.zeros(3, np.float32); N.check(lib.sacb_get_losses(h, 0, N.ptr(losses)))
    print(f"steps {(blk + 1) * 2000:6d}: {ms.value / 2000 * 1e3:7.2f} us/step  fine {st.n_fine:6d}  exact fallbacks {st.n_exact_fallbacks:4d}  total {st.total_f32:.1f}  losses {losses.tolist()}", flush=True)
    assert np.all(np.isfinite(losses))
