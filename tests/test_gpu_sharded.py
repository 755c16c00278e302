"""GPU tests of the two sharded modes on ONE device: a population of independent agents in one
handle must equal the same agents run one by one, and the data-parallel step over two logical shards must equal the
single-agent step on the concatenated batch."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import sac_oracle_np as O
from tests.golden import cases
from tests.util import batch_of, make_agent, net_params, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hw():
    import humanoid_walking_with_sac_b200 as hw
    return hw


def _push_all(agent, b):
    agent.replay_buffer.push_many(b["s"], b["a"], b["r"], b["s2"], b["d"])


def test_population_equals_independent_agents(hw):
    """BASELINE.json configs[4] in miniature: 3 agents with distinct weights / data / eps in one handle (agent = grid
    dimension and 4th TMA coordinate) against three single-agent handles.  No communication exists to test."""
    N = hw._native
    lib = N.lib()
    case = cases.UPDATE_CASES["tiny_m2"]
    B, n_agents = case["batch"], 3
    singles, batches = [], []
    for a in range(n_agents):
        agent, _ = make_agent(hw, dict(case, seed=case["seed"] + 10 * a), math="bf16x3", capacity=64)
        b = batch_of(dict(case, seed=case["seed"] + 10 * a), 0)
        _push_all(agent, b)
        singles.append(agent)
        batches.append(b)
    cfg = N.default_config()
    c0 = singles[0]._cfg
    for f, _ in N.Config._fields_:
        if f != "reserved":
            setattr(cfg, f, getattr(c0, f))
    cfg.n_agents = n_agents
    h = N.create(cfg)
    try:
        nets = ("policy", "q1", "q2", "q1_target", "q2_target")
        for a, agent in enumerate(singles):
            for net_id, net in enumerate(nets):
                for t, (_, p) in enumerate(getattr(agent, net).named_parameters()):
                    host = N.f32(p.detach().cpu().numpy())
                    N.check(lib.sacb_import_tensor(h, a, net_id, N.SLOT_PARAM, t, N.ptr(host), host.size))
            b = batches[a]
            s, ac, r, s2, d = (N.f32(b[k]) for k in ("s", "a", "r", "s2", "d"))
            N.check(lib.sacb_push(h, a, N.ptr(s), N.ptr(ac), N.ptr(r), N.ptr(s2), N.ptr(d), B))
        idx = np.ascontiguousarray(np.tile(np.arange(B, dtype=np.int64), (n_agents, 1)))
        e_next = np.ascontiguousarray(np.stack([b["eps_next"] for b in batches]), np.float32)
        e_cur = np.ascontiguousarray(np.stack([b["eps_cur"] for b in batches]), np.float32)
        N.check(lib.sacb_update(h, B, N.ptr(idx, ctypes.c_int64), N.ptr(e_next), N.ptr(e_cur), None, N.NO_LOSS_READBACK))
        N.check(lib.sacb_synchronize(h))
        for a, agent in enumerate(singles):
            got = agent.update_parameters(B, idx=np.arange(B), eps=(batches[a]["eps_next"], batches[a]["eps_cur"]))
            losses = np.zeros(3, np.float32)
            N.check(lib.sacb_get_losses(h, a, N.ptr(losses)))
            np.testing.assert_array_equal(losses, np.array([got["q1_loss"], got["q2_loss"], got["policy_loss"]], np.float32))
            for net_id, net in enumerate(nets):
                for t, (nm, p) in enumerate(getattr(agent, net).named_parameters()):
                    out = np.empty(p.numel(), np.float32)
                    N.check(lib.sacb_export_tensor(h, a, net_id, N.SLOT_PARAM, t, N.ptr(out), out.size))
                    np.testing.assert_array_equal(out.reshape(p.shape), p.detach().cpu().numpy(), err_msg=f"agent {a} {net}.{nm}")
        # population select_action: row i of the batch acts with agent i's policy, one chain of launches
        rng = np.random.RandomState(5)
        obs = rng.standard_normal((n_agents, case["obs"])).astype(np.float32)
        eps = rng.standard_normal((n_agents, case["act"])).astype(np.float32)
        for evaluate, e in ((1, None), (0, eps)):
            act = np.empty((n_agents, case["act"]), np.float32)
            N.check(lib.sacb_select_action_batch(h, N.ptr(obs), evaluate, N.ptr(e), N.ptr(act)))
            for a, agent in enumerate(singles):
                ref = agent.select_action(obs[a], evaluate=bool(evaluate), eps=None if e is None else e[a:a + 1])
                np.testing.assert_array_equal(act[a], ref, err_msg=f"agent {a} evaluate={evaluate}")
    finally:
        lib.sacb_destroy(h)


@pytest.mark.parametrize("name", ["tiny_m2", "c1_bipedal_m1"])
def test_data_parallel_two_shards_equal_full_batch(hw, name):
    """BASELINE.json configs[3] on one GPU: two replicas, half the batch each, gradient slabs averaged by hand exactly as
    the all-reduce would; result == the single-agent update on the whole batch up to fp32 reassociation of the batch sum."""
    N = hw._native
    lib = N.lib()
    case = cases.UPDATE_CASES[name]
    B = case["batch"]
    b = batch_of(case, 0)
    full, _ = make_agent(hw, case, math="bf16x3", capacity=1024)
    _push_all(full, b)
    ref = full.update_parameters(B, idx=np.arange(B), eps=(b["eps_next"], b["eps_cur"]))
    shards = []
    for r in range(2):
        ag, _ = make_agent(hw, case, math="bf16x3", capacity=1024)
        _push_all(ag, b)
        shards.append(hw.distributed.DataParallelSAC(ag))
    half = B // 2
    rows = [np.arange(0, half, dtype=np.int64), np.arange(half, B, dtype=np.int64)]
    for phase in (0, 1):
        for r, dp in enumerate(shards):
            ix = rows[r]
            e = (N.f32(b["eps_next"][ix]), N.f32(b["eps_cur"][ix]))
            N.check(lib.sacb_dp_backward(dp.agent._h, phase, half, N.ptr(ix, ctypes.c_int64) if phase == 0 else None, N.ptr(e[0]), N.ptr(e[1])))
            dp.agent.synchronize()
        for s0, s1 in zip(shards[0].gradient_slabs(phase), shards[1].gradient_slabs(phase)):      # what all-reduce(mean) does
            m = (s0 + s1) / 2
            s0.copy_(m); s1.copy_(m)
        torch.cuda.synchronize()
        for dp in shards:
            N.check(lib.sacb_dp_apply(dp.agent._h, phase))
            dp.agent.synchronize()
    lr = 3e-4
    for net in ("policy", "q1", "q2", "q1_target", "q2_target"):
        want = net_params(full, net)
        for r in range(2):
            mine = net_params(shards[r].agent, net)
            for nm in want:
                # identical gradients up to reassociation; Adam's first step is sign-like, so a gradient within rounding of
                # zero may move by +lr in one run and -lr in the other: allow a handful of such elements
                bad = np.mean(np.abs(mine[nm] - want[nm]) > 0.02 * lr)
                assert bad < 5e-3, (net, nm, r, bad)
    np.testing.assert_array_equal(net_params(shards[0].agent, "q1")["fc1.weight"], net_params(shards[1].agent, "q1")["fc1.weight"])
    a_full, a_dp = float(torch.as_tensor(full.alpha).reshape(-1)[0]), float(torch.as_tensor(shards[0].agent.alpha).reshape(-1)[0])
    assert abs(a_full - a_dp) <= 1e-6 * abs(a_full)
    l0 = np.zeros(3, np.float32); l1 = np.zeros(3, np.float32)
    N.check(lib.sacb_get_losses(shards[0].agent._h, 0, N.ptr(l0))); N.check(lib.sacb_get_losses(shards[1].agent._h, 0, N.ptr(l1)))
    np.testing.assert_allclose((l0 + l1) / 2, [ref["q1_loss"], ref["q2_loss"], ref["policy_loss"]], rtol=2e-5)


def test_data_parallel_wrapper_world1_matches_plain_update(hw):
    """DataParallelSAC without a process group (world = 1) is the plain update: backward/export + apply kernels vs the fused epilogues."""
    case = cases.UPDATE_CASES["tiny_m1"]
    B = case["batch"]
    b = batch_of(case, 0)
    plain, _ = make_agent(hw, case, math="bf16x3", capacity=256)
    dp_agent, _ = make_agent(hw, case, math="bf16x3", capacity=256)
    _push_all(plain, b); _push_all(dp_agent, b)
    ref = plain.update_parameters(B, idx=np.arange(B), eps=(b["eps_next"], b["eps_cur"]))
    got = hw.distributed.DataParallelSAC(dp_agent).update_parameters(B, idx=np.arange(B), eps=(b["eps_next"], b["eps_cur"]))
    for k in ref:
        assert abs(got[k] - ref[k]) <= 1e-6 * abs(ref[k]) + 1e-7
    for net in ("policy", "q1", "q2", "q1_target"):
        w, m = net_params(plain, net), net_params(dp_agent, net)
        for nm in w:
            assert np.mean(np.abs(w[nm] - m[nm]) > 0.02 * 3e-4) < 5e-3, (net, nm)
    assert dp_agent.policy_optimizer.state_dict()["state"][0]["step"] == 1
