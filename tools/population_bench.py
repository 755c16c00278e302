"""Population mode (BASELINE.json configs[4]): A independent C2 agents in one handle, agent-updates/s on one GPU.
usage: population_bench.py [n_agents] [steps] [launch]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import humanoid_walking_with_sac_b200 as hw
N = hw._native
lib = N.lib()
A = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
launch = sys.argv[3] if len(sys.argv) > 3 else "staged"
OBS, ACT, HID, NH, B, CAP = 348, 17, 512, 3, 256, 2048
cfg = N.default_config()
cfg.obs_dim, cfg.act_dim, cfg.hidden_dim, cfg.n_hidden = OBS, ACT, HID, NH
cfg.capacity, cfg.max_batch, cfg.n_agents = CAP, B, A
cfg.launch_mode = N.LAUNCH_PERSISTENT if launch == "persistent" else N.LAUNCH_STAGED
h = N.create(cfg)
rng = np.random.RandomState(0)
shapes = {"policy": [(HID, OBS), (HID,), (HID, HID), (HID,), (HID, HID), (HID,), (ACT, HID), (ACT,), (ACT, HID), (ACT,)],
          "q": [(HID, OBS + ACT), (HID,), (HID, HID), (HID,), (HID, HID), (HID,), (1, HID), (1,)]}
for a in range(A):
    for net in range(5):
        for t, shp in enumerate(shapes["policy" if net == 0 else "q"]):
            w = (rng.uniform(-1, 1, shp) * (np.sqrt(6.0 / sum(shp)) if len(shp) == 2 else 0.0)).astype(np.float32)
            if net == 0 and t in (6, 8):
                w *= 0.25
            N.check(lib.sacb_import_tensor(h, a, net, N.SLOT_PARAM, t, N.ptr(w), w.size))
    s = rng.standard_normal((CAP, OBS)).astype(np.float32); s2 = rng.standard_normal((CAP, OBS)).astype(np.float32)
    ac = rng.uniform(-0.4, 0.4, (CAP, ACT)).astype(np.float32); r = rng.standard_normal(CAP).astype(np.float32)
    d = (rng.uniform(size=CAP) < 0.01).astype(np.float32)
    N.check(lib.sacb_push(h, a, N.ptr(s), N.ptr(ac), N.ptr(r), N.ptr(s2), N.ptr(d), CAP))
n_sets = 8
idx = np.ascontiguousarray(rng.randint(0, CAP, size=(n_sets, A, B)), np.int64)
N.check(lib.sacb_stage_indices(h, N.ptr(idx, ctypes.c_int64), B, n_sets))
for _ in range(3):
    N.check(lib.sacb_update(h, B, None, None, None, None, N.NO_LOSS_READBACK))
N.check(lib.sacb_synchronize(h))
ms = ctypes.c_float()
N.check(lib.sacb_timer_start(h))
for _ in range(steps):
    N.check(lib.sacb_update(h, B, None, None, None, None, N.NO_LOSS_READBACK))
N.check(lib.sacb_timer_stop(h, ctypes.byref(ms)))
losses = np.zeros(3, np.float32)
N.check(lib.sacb_get_losses(h, A - 1, N.ptr(losses)))
per_step = ms.value / steps
print(f"POPULATION agents={A} launch={launch} ms_per_population_step={per_step:.3f} agent_updates_per_s={A / per_step * 1e3:.0f} "
      f"algorithmic_TFLOPs={A * 5.382e9 / (per_step * 1e-3) / 1e12:.1f} last_agent_losses={losses.tolist()}")
lib.sacb_destroy(h)
