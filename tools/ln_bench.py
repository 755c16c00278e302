"""Update time of the opt-in LayerNorm variant against the reference architecture (C2 shape, B = 256).  usage: ln_bench.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import humanoid_walking_with_sac_b200 as hw
from oracle import sac_oracle_np as O
N = hw._native
hw.use_networks("model2")
for ln in (False, True):
    torch.manual_seed(0)
    agent = hw.SAC(348, 17, hidden_dim=512, device="cuda", max_batch=256, capacity=1024, seed=1, layer_norm=ln)
    b = O.make_batch(348, 17, 256, seed=3)
    agent.update_from_batch(b)
    best = 1e9
    for _ in range(5):
        ms = ctypes.c_float()
        N.check(N.lib().sacb_time_update(agent._h, 256, 300, ctypes.byref(ms)))
        best = min(best, ms.value)
    print(f"LN_BENCH layer_norm={ln} update_ms={best:.4f} stages={agent.stats()['n_stages']} select_action_us=", end="")
    import time
    s = b["s"][0]
    agent.select_action(s)
    t0 = time.perf_counter()
    for _ in range(200):
        agent.select_action(s)
    print(f"{(time.perf_counter() - t0) / 200 * 1e6:.1f}")
