"""torchrun check of the data-parallel mode over NCCL: G ranks x B/G rows == one GPU x B rows (run: torchrun --nproc-per-node G tools/dp_check.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import humanoid_walking_with_sac_b200 as hw
from tests.golden import cases
from tests.util import batch_of, make_agent, net_params

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
case = dict(cases.UPDATE_CASES["c1_bipedal_m1"])
B = case["batch"]
b = batch_of(case, 0)
hw.use_networks("model1")
import tests.util as U
_orig = hw.SAC
agent, _ = make_agent(type("H", (), {"SAC": lambda *a, **k: _orig(*a, **{**k, "device": f"cuda:{local}"}), "use_networks": hw.use_networks}), case, math="bf16x3", capacity=1024)
agent.replay_buffer.push_many(b["s"], b["a"], b["r"], b["s2"], b["d"])
rows = np.arange(rank * (B // world), (rank + 1) * (B // world), dtype=np.int64)
dp = hw.distributed.DataParallelSAC(agent)
out = dp.update_parameters(len(rows), idx=rows, eps=(b["eps_next"][rows], b["eps_cur"][rows]))
if rank == 0:
    full, _ = make_agent(type("H", (), {"SAC": lambda *a, **k: _orig(*a, **{**k, "device": f"cuda:{local}"}), "use_networks": hw.use_networks}), case, math="bf16x3", capacity=1024)
    full.replay_buffer.push_many(b["s"], b["a"], b["r"], b["s2"], b["d"])
    ref = full.update_parameters(B, idx=np.arange(B), eps=(b["eps_next"], b["eps_cur"]))
    worst = 0.0
    for net in ("policy", "q1", "q2", "q1_target"):
        w, m = net_params(full, net), net_params(agent, net)
        for nm in w:
            worst = max(worst, float(np.mean(np.abs(w[nm] - m[nm]) > 0.02 * 3e-4)))
    print("DP_CHECK world", world, "losses dp", out, "ref", ref, "worst frac of weights off by > 0.02 lr:", worst)
    assert worst < 5e-3
    for k in ref:
        assert abs(out[k] - ref[k]) <= 2e-5 * abs(ref[k]) + 1e-7, (k, out[k], ref[k])
    print("DP_CHECK OK")
dist.barrier()
dist.destroy_process_group()
