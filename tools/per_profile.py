"""PER sampler micro-benchmark: N-row prioritized buffer, `iters` device-uniform sample(256) calls (+ write-back), CUDA-event time per call.
usage: per_profile.py [N] [iters] [writeback 0|1]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import humanoid_walking_with_sac_b200 as hw
N_ = hw._native
lib = N_.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
wb = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rng = np.random.RandomState(0)
buf = hw.PrioritizedReplayBuffer(n)
chunk = 100_000
s = rng.standard_normal((chunk, 4)).astype(np.float32); a = rng.standard_normal((chunk, 2)).astype(np.float32)
for _ in range(n // chunk):
    buf.push_many(s, a, s[:, 0], s, np.zeros(chunk, np.float32))
buf.set_priorities((np.abs(rng.standard_normal(n)) + 1e-6).astype(np.float32))
h = buf._h
def call():
    N_.check(lib.sacb_per_sample(h, 0, None, 256, None, None, None, None, None, None, None))
    if wb:
        N_.check(lib.sacb_per_update_from_td(h, 0, 256))
for _ in range(5):
    call()
ms = ctypes.c_float()
N_.check(lib.sacb_timer_start(h))
for _ in range(iters):
    call()
N_.check(lib.sacb_timer_stop(h, ctypes.byref(ms)))
print(f"N={n} sample(256){' + write-back' if wb else ''}: {ms.value / iters * 1e3:.2f} us per call, {256 * iters / (ms.value * 1e-3) / 1e6:.2f} M samples/s")
