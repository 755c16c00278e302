"""PER sample(256) at N = 1 M on the adversarial priority sets of SURVEY H6.4 (many "fine" probabilities): average time per call and how
often the exact pass runs."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import humanoid_walking_with_sac_b200 as hw
from tests.golden import cases
N_ = hw._native; lib = N_.lib()
n, iters = 1_000_000, 300
rows = None
for dist in ("halfnormal", "floor1pct", "lognormal3", "fresh"):
    buf = hw.PrioritizedReplayBuffer(n)
    buf.push_many(np.zeros((1, 2)), np.zeros((1, 1)), np.zeros(1), np.zeros((1, 2)), np.zeros(1))
    rows = np.zeros((n - 1, int(lib.sacb_row_floats(buf._h))), np.float32)
    N_.check(lib.sacb_push_rows(buf._h, 0, N_.ptr(rows), n - 1))
    buf.set_priorities(cases.per_priorities(dict(n=n, capacity=n, seed=31, dist=dist)))
    h = buf._h
    for _ in range(5):
        N_.check(lib.sacb_per_sample(h, 0, None, 256, None, None, None, None, None, None, None))
    ms = ctypes.c_float()
    N_.check(lib.sacb_timer_start(h))
    for _ in range(iters):
        N_.check(lib.sacb_per_sample(h, 0, None, 256, None, None, None, None, None, None, None))
    N_.check(lib.sacb_timer_stop(h, ctypes.byref(ms)))
    st = buf._stats()
    print(f"{dist:11s}: {ms.value / iters * 1e3:8.1f} us per sample(256)   fine elements {st.n_fine:7d}   exact passes {st.n_exact_fallbacks} / {iters + 5} calls")
    del buf
