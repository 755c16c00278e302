"""Large-batch data-parallel mode (BASELINE.json configs[3]): global batch split over the ranks, two gradient all-reduces per step
over NCCL.  run: torchrun --nproc-per-node G --master-addr 127.0.0.1 tools/dp_bench.py [global_batch] [steps]   (G = 1 works too)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import humanoid_walking_with_sac_b200 as hw

N = hw._native
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
GB = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
OBS, ACT, HID = 348, 17, 512
BL, CAP = GB // world, 65536
hw.use_networks("model2")
torch.manual_seed(0)                      # identical replicas
agent = hw.SAC(OBS, ACT, hidden_dim=HID, device=f"cuda:{local}", capacity=CAP, max_batch=BL, math="bf16x3", seed=100 + rank)
with torch.no_grad():
    agent.policy.mean.weight.mul_(0.25); agent.policy.log_std.weight.mul_(0.25)
rng = np.random.RandomState(rank)
agent.replay_buffer.push_many(rng.standard_normal((CAP, OBS)).astype(np.float32), rng.uniform(-0.4, 0.4, (CAP, ACT)).astype(np.float32),
                              rng.standard_normal(CAP).astype(np.float32), rng.standard_normal((CAP, OBS)).astype(np.float32),
                              (rng.uniform(size=CAP) < 0.01).astype(np.float32))
dp = hw.distributed.DataParallelSAC(agent, exchange=os.environ.get('SACB_DP_EXCHANGE', 'peer'))
idx = rng.randint(0, CAP, BL).astype(np.int64)
out = dp.update_parameters(BL, idx=idx)                 # stages this rank's rows on the device; device-drawn eps
for _ in range(3):
    dp.update_parameters(BL, staged=True, sync=False)
agent.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
ms = ctypes.c_float()
N.check(N.lib().sacb_timer_start(agent._h))
for _ in range(steps):
    dp.update_parameters(BL, staged=True, sync=False)
N.check(N.lib().sacb_timer_stop(agent._h, ctypes.byref(ms)))
t = torch.tensor([ms.value], device=f"cuda:{local}")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
per = float(t.item()) / steps
last = dp.update_parameters(BL, staged=True)
if rank == 0:
    print(f"DP_BENCH exchange={dp.mode} world={world} global_batch={GB} local_batch={BL} ms_per_step={per:.3f} updates_per_s={1e3 / per:.1f} "
          f"transitions_per_s={GB / per * 1e3:.3e} algorithmic_TFLOPs={5.382e9 * GB / 256 / (per * 1e-3) / 1e12:.1f} losses={last}")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
