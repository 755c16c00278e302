"""End-to-end step over the UNIFORM ring (the reference's default buffer, sac_imp.py:52): push of one transition + update_parameters(256)
with `random.sample`-identical host index draws, losses read back every step.  C2 nets, 1 M-row ring."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import humanoid_walking_with_sac_b200 as hw
hw.use_networks("model2")
torch.manual_seed(0)
agent = hw.SAC(bench.OBS, bench.ACT, hidden_dim=bench.HID, device="cuda", capacity=bench.CAPACITY, max_batch=bench.B, seed=0)
with torch.no_grad():
    agent.policy.mean.weight.mul_(0.25); agent.policy.log_std.weight.mul_(0.25)
s, a, r, s2, d = bench.synth_transitions(125_000, 0)
for _ in range(8):
    agent.replay_buffer.push_many(s, a, r, s2, d)
n = 400
for i in range(20):
    agent.replay_buffer.push(s[i], a[i], float(r[i]), s2[i], bool(d[i])); agent.update_parameters(bench.B)
agent.synchronize()
t0 = time.perf_counter()
for i in range(n):
    agent.replay_buffer.push(s[i], a[i], float(r[i]), s2[i], bool(d[i]))
    out = agent.update_parameters(bench.B)
dt = (time.perf_counter() - t0) / n
print(f"uniform-ring e2e: {dt * 1e6:.1f} us per step = {1 / dt:.0f} updates/s", out)
