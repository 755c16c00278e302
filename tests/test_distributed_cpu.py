"""Host-side logic of the sharded modes on CPU: world_size-2 gloo process group.

The data path of the single-agent update does not shard (replicas only); what is exercised here is the plumbing the
multi-GPU modes add on the host: agent partitioning of a population (no communication) and the gradient averaging of
the data-parallel mode (mean of equal-sized means == global mean)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import humanoid_walking_with_sac_b200 as hw
    D = hw.distributed
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(100 + rank)
        rows = rng.standard_normal((8, 5)).astype(np.float32)          # this rank's shard of a global batch of 16 rows
        local_mean = torch.from_numpy(rows.mean(axis=0))
        scal = torch.tensor([float(rank + 1)])
        D.allreduce_mean_([local_mean, scal])
        owned = list(D.partition_agents(11, world, rank))
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), mean=local_mean.numpy(), rows=rows, scal=scal.numpy(), owned=np.array(owned))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gradient_average_and_agent_partition(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"r{k}.npz") for k in range(world)]
    global_mean = np.concatenate([r[0]["rows"], r[1]["rows"]]).mean(axis=0)
    for k in range(world):
        np.testing.assert_allclose(r[k]["mean"], global_mean, rtol=1e-6, atol=1e-7)     # mean of equal-sized means
        np.testing.assert_allclose(r[k]["scal"], [1.5])
    owned = np.concatenate([r[0]["owned"], r[1]["owned"]])
    assert sorted(owned.tolist()) == list(range(11))                                     # every agent owned exactly once
    assert abs(len(r[0]["owned"]) - len(r[1]["owned"])) <= 1


def test_partition_agents_covers_population():
    import humanoid_walking_with_sac_b200 as hw
    for n, w in ((1024, 8), (7, 3), (3, 8), (1, 1)):
        parts = [list(hw.distributed.partition_agents(n, w, r)) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_allreduce_mean_is_identity_without_a_group():
    import humanoid_walking_with_sac_b200 as hw
    t = torch.arange(4.0)
    hw.distributed.allreduce_mean_([t])
    assert t.tolist() == [0.0, 1.0, 2.0, 3.0]
