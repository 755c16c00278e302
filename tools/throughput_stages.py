"""Per-stage device time of the update program in the two throughput modes (large batch / population), for tile-shape work.
usage: throughput_stages.py [n_agents] [B]     (SACB_TRACE=1 adds the per-task breakdown on stderr)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import humanoid_walking_with_sac_b200 as hw
N = hw._native
lib = N.lib()
A = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
OBS, ACT, HID, NH = 348, 17, 512, 3
cfg = N.default_config()
cfg.obs_dim, cfg.act_dim, cfg.hidden_dim, cfg.n_hidden = OBS, ACT, HID, NH
cfg.capacity, cfg.max_batch, cfg.n_agents = 1024, B, A
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
us = (ctypes.c_float * 64)()
n = lib.sacb_time_stages(h, B, us, 64)
N.check(min(n, 0))
ms = ctypes.c_float()
N.check(lib.sacb_time_update(h, B, 10, ctypes.byref(ms)))
tot = sum(us[i] for i in range(n))
print(f"THROUGHPUT_STAGES agents={A} B={B} stages={n} sum_us={tot:.1f} update_ms={ms.value:.3f} algorithmic_TFLOPs={A * 5.382e9 * B / 256 / (ms.value * 1e-3) / 1e12:.1f}")
print("stage_us", " ".join(f"{us[i]:.1f}" for i in range(n)))
lib.sacb_destroy(h)
