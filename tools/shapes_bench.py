"""Update-program time (sacb_time_update: CUDA events over graph replays, device-drawn eps, no replay traffic) for the BASELINE.json shapes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import humanoid_walking_with_sac_b200 as hw
N = hw._native
SHAPES = [("C1 BipedalWalker m1 24/4/2x256", "model1", 24, 4, 256, 0.582e9), ("Humanoid m1 348/17/2x256 (trainer default)", "model1", 348, 17, 256, 1.080e9),
          ("C2 Humanoid m2 348/17/3x512", "model2", 348, 17, 512, 5.382e9), ("C3 NAO m2 661/23/3x512", "model2", 661, 23, 512, 6.313e9)]
for name, nets, obs, act, hid, flop in SHAPES:
    hw.use_networks(nets)
    torch.manual_seed(0)
    agent = hw.SAC(obs, act, hidden_dim=hid, device="cuda", capacity=4096, max_batch=256)
    ms = ctypes.c_float()
    N.check(N.lib().sacb_time_update(agent._h, 256, 200, ctypes.byref(ms)))
    st = agent.stats()
    print(f"{name}: {ms.value * 1e3:.1f} us per update = {1e3 / ms.value:.0f} updates/s, {st['n_stages']} stages, {flop / (ms.value * 1e-3) / 1e12:.1f} algorithmic TFLOP/s")
    del agent
