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
