"""select_action latency through the public API (B = 1, Humanoid 3x512 policy).  usage: act_latency.py [iters]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import humanoid_walking_with_sac_b200 as hw
hw.use_networks("model2")
agent = hw.SAC(348, 17, hidden_dim=512, device="cuda", capacity=1024, max_batch=256)
obs = np.random.RandomState(0).standard_normal(348)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for ev in (True, False):
    for _ in range(50): agent.select_action(obs, evaluate=ev)
    t0 = time.perf_counter()
    for _ in range(n): a = agent.select_action(obs, evaluate=ev)
    dt = (time.perf_counter() - t0) / n
    print(f"select_action(evaluate={ev}): {dt * 1e6:.1f} us per call", a[:3])
