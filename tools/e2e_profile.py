"""Host-side breakdown of the e2e step (push + update_parameters through the public API)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import humanoid_walking_with_sac_b200 as hw

agent = bench.build_agent(hw, 0, "staged", "bf16x3", seed=0)
s1, a1, r1, s21, d1 = bench.synth_transitions(600, 1)
for i in range(20):
    agent.replay_buffer.push(s1[i], a1[i], float(r1[i]), s21[i], bool(d1[i])); agent.update_parameters(256)
agent.synchronize()
n = 300
t_push = t_upd = 0.0
for i in range(n):
    t0 = time.perf_counter(); agent.replay_buffer.push(s1[i], a1[i], float(r1[i]), s21[i], bool(d1[i])); t1 = time.perf_counter()
    agent.update_parameters(256); t2 = time.perf_counter()
    t_push += t1 - t0; t_upd += t2 - t1
print(f"push {t_push/n*1e6:.1f} us  update_parameters {t_upd/n*1e6:.1f} us  total {(t_push+t_upd)/n*1e6:.1f} us")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(n):
    agent.replay_buffer.push(s1[i], a1[i], float(r1[i]), s21[i], bool(d1[i])); agent.update_parameters(256)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

# ---- the same sequence call by call (what update_parameters does for the prioritized buffer), host clock around every C call
import ctypes
from humanoid_walking_with_sac_b200 import _native as N
lib, h, buf = N.lib(), agent._h, agent.replay_buffer
names = ["push(py)", "random(py)", "push_rows", "per_sample", "update+loss sync", "write-back", "dict(py)"]
acc = np.zeros(len(names))
losses = np.zeros(3, np.float32)
for i in range(n):
    t = [time.perf_counter()]
    buf.push(s1[i], a1[i], float(r1[i]), s21[i], bool(d1[i])); t.append(time.perf_counter())
    uu = np.ascontiguousarray(np.random.random_sample(256), np.float64); t.append(time.perf_counter())
    buf._flush(); t.append(time.perf_counter())
    N.check(lib.sacb_per_sample(h, 0, N.ptr(uu, ctypes.c_double), 256, None, None, None, None, None, None, None)); t.append(time.perf_counter())
    N.check(lib.sacb_update(h, 256, None, None, None, N.ptr(losses), N.USE_LAST_SAMPLE)); t.append(time.perf_counter())
    N.check(lib.sacb_per_update_from_td(h, 0, 256)); t.append(time.perf_counter())
    out = {"q1_loss": float(losses[0]), "q2_loss": float(losses[1]), "policy_loss": float(losses[2])}; t.append(time.perf_counter())
    acc += np.diff(t)
print("per call, us: " + "  ".join(f"{k} {v / n * 1e6:.1f}" for k, v in zip(names, acc)) + f"  total {acc.sum() / n * 1e6:.1f}")
# device time of the same step: sample -> update -> write-back enqueued back to back, no host work in between
ms = ctypes.c_float()
N.check(lib.sacb_timer_start(h))
for i in range(n):
    N.check(lib.sacb_per_sample(h, 0, None, 256, None, None, None, None, None, None, None))
    N.check(lib.sacb_update(h, 256, None, None, None, None, N.USE_LAST_SAMPLE | N.NO_LOSS_READBACK))
    N.check(lib.sacb_per_update_from_td(h, 0, 256))
N.check(lib.sacb_timer_stop(h, ctypes.byref(ms)))
print(f"device time of sample -> update -> write-back back to back: {ms.value / n * 1e3:.1f} us per step")
