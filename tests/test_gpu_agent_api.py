"""GPU tests of the rest of the SAC surface: select_action, update_parameters over the replay ring, checkpoints."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import sac_oracle_np as O
from tests.golden import cases
from tests.util import batch_of, make_agent, net_params, relerr, relu_hint

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def hw():
    import humanoid_walking_with_sac_b200 as hw
    return hw


@pytest.mark.parametrize("name", ["tiny_m1", "tiny_m2", "c2_humanoid_m2"])
def test_select_action_matches_reference(hw, name):
    case = cases.UPDATE_CASES[name]
    g = np.load(os.path.join(GOLD, f"update_{name}.npz"))
    agent, st = make_agent(hw, case, math="fp32")
    for step in range(case["steps"]):
        b = batch_of(case, step)
        agent.update_from_batch(b, eps=(b["eps_next"], b["eps_cur"]))
        O.update_parameters(st, b, relu_hint=relu_hint(agent, case))      # oracle follows the device's ReLU tie-breaks
    obs_vec = np.random.RandomState(77 + case["seed"]).standard_normal(case["obs"]).astype(np.float32)
    eps_vec = np.random.RandomState(78 + case["seed"]).standard_normal((1, case["act"])).astype(np.float32)
    # against the oracle that took the same steps (tight) and against the live reference's golden vector (a ReLU tie or an
    # Adam sign tie in one of the preceding updates moves single weights by 2*lr: looser)
    np.testing.assert_allclose(agent.select_action(obs_vec, evaluate=True), O.select_action(st, obs_vec, evaluate=True), rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(agent.select_action(obs_vec, eps=eps_vec), O.select_action(st, obs_vec, eps=eps_vec[0]), rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(agent.select_action(obs_vec, evaluate=True), g["select/eval"], rtol=1e-2, atol=2e-4)
    np.testing.assert_allclose(agent.select_action(obs_vec, eps=eps_vec), g["select/sample"], rtol=1e-2, atol=2e-4)
    a = agent.select_action(obs_vec)                      # production draw: inside the action bounds
    assert a.shape == (case["act"],) and np.all(np.abs(a) <= 0.4 + 1e-6)


def test_update_parameters_over_uniform_ring(hw):
    """The trainer's path (trainer.py:194-204): push transitions, update_parameters(B) draws with random.sample."""
    case = cases.UPDATE_CASES["tiny_m1"]
    agent, st = make_agent(hw, case, math="fp32", capacity=64)
    rng = np.random.RandomState(3)
    n = 100                                                # > capacity: ring wraps (deque eviction)
    S, A = rng.standard_normal((n, case["obs"])), rng.uniform(-0.4, 0.4, (n, case["act"]))
    R, S2, D = rng.standard_normal(n), rng.standard_normal((n, case["obs"])), rng.uniform(size=n) < 0.1
    for i in range(n):
        agent.replay_buffer.push(S[i], A[i].astype(np.float32), R[i], S2[i], bool(D[i]))
    assert len(agent.replay_buffer) == 64
    B = case["batch"]
    random.seed(11)
    pos = random.sample(range(64), B)
    logical = np.arange(n - 64, n)[pos]                    # j-th oldest of the surviving 64
    eps_next = rng.standard_normal((B, case["act"])).astype(np.float32)
    eps_cur = rng.standard_normal((B, case["act"])).astype(np.float32)
    batch = dict(s=S[logical].astype(np.float32), a=A[logical].astype(np.float32), r=R[logical].astype(np.float32),
                 s2=S2[logical].astype(np.float32), d=D[logical].astype(np.float32), eps_next=eps_next, eps_cur=eps_cur)
    ref = O.update_parameters(st, batch)
    random.seed(11)
    got = agent.update_parameters(B, eps=(eps_next, eps_cur))
    for k in ref:
        assert abs(got[k] - ref[k]) <= 2e-4 * abs(ref[k]) + 1e-6, (k, got[k], ref[k])
    with pytest.raises(ValueError):
        agent.update_parameters(65)
    out = agent.update_parameters(B)                       # production mode (device eps)
    assert set(out) == {"q1_loss", "q2_loss", "policy_loss"} and all(np.isfinite(v) for v in out.values())


def test_update_parameters_over_prioritized_ring(hw):
    case = cases.UPDATE_CASES["tiny_m2"]
    agent, st = make_agent(hw, case, math="fp32", capacity=512, replay="per")
    rng = np.random.RandomState(4)
    n = 300
    S, A = rng.standard_normal((n, case["obs"])).astype(np.float32), rng.uniform(-0.4, 0.4, (n, case["act"])).astype(np.float32)
    R, S2, D = rng.standard_normal(n).astype(np.float32), rng.standard_normal((n, case["obs"])).astype(np.float32), (rng.uniform(size=n) < 0.1)
    agent.replay_buffer.push_many(S, A, R, S2, D)
    pri = (np.abs(rng.standard_normal(n)) + 1e-6).astype(np.float32)
    from oracle import per_oracle as PO
    pa = PO.pow_alpha(pri)
    full_p, full_pa = np.zeros(512, np.float32), np.zeros(512, np.float32)
    full_p[:n], full_pa[:n] = pri, pa
    agent.replay_buffer.set_priorities(full_p, full_pa)
    B = case["batch"]
    u = rng.random_sample(B)
    idx, _ = PO.sample(pa, u, PO.beta(1))
    eps_next = rng.standard_normal((B, case["act"])).astype(np.float32)
    eps_cur = rng.standard_normal((B, case["act"])).astype(np.float32)
    batch = dict(s=S[idx], a=A[idx], r=R[idx], s2=S2[idx], d=D[idx].astype(np.float32), eps_next=eps_next, eps_cur=eps_cur)
    ref = O.update_parameters(st, batch)
    got = agent.update_parameters(B, eps=(eps_next, eps_cur), u=u)
    for k in ref:
        assert abs(got[k] - ref[k]) <= 2e-4 * abs(ref[k]) + 1e-6, (k, got[k], ref[k])


def test_checkpoint_round_trip_and_layout(hw, tmp_path):
    case = cases.UPDATE_CASES["tiny_m1"]
    agent, st = make_agent(hw, case, math="fp32", capacity=128)
    for step in range(2):
        b = batch_of(case, step)
        agent.update_from_batch(b, eps=(b["eps_next"], b["eps_cur"]))
    for i in range(20):
        agent.replay_buffer.push(np.full(case["obs"], i), np.full(case["act"], 0.1), float(i), np.full(case["obs"], i + 1), i == 7)
    p1, p2 = str(tmp_path / "model.pt"), str(tmp_path / "ckpt.pt")
    agent.save(p1)
    agent.save_checkpoint(p2, episode=5, total_steps=99)
    ck = torch.load(p2, weights_only=False)
    # reference key set (sac_imp.py:178-199) and torch-Adam state layout
    assert set(ck) == {"episode", "total_steps", "policy_state_dict", "q1_state_dict", "q2_state_dict", "q1_target_state_dict",
                       "q2_target_state_dict", "policy_optimizer_state_dict", "q1_optimizer_state_dict", "q2_optimizer_state_dict",
                       "alpha", "log_alpha", "alpha_optimizer_state_dict", "replay_buffer"}
    osd = ck["q1_optimizer_state_dict"]
    assert set(osd) == {"state", "param_groups"} and set(osd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert osd["param_groups"][0]["params"] == list(range(6)) and float(osd["state"][3]["step"]) == 2.0
    assert len(ck["replay_buffer"]) == 20 and ck["replay_buffer"][7][4] is True
    # a torch.optim.Adam built on same-shaped parameters accepts the state dict unchanged
    ref_q = torch.nn.ModuleList([torch.nn.Linear(case["obs"] + case["act"], case["hidden"]), torch.nn.Linear(case["hidden"], case["hidden"]),
                                 torch.nn.Linear(case["hidden"], 1)])
    torch.optim.Adam(ref_q.parameters(), lr=3e-4).load_state_dict({"state": {k: {kk: vv.cpu() for kk, vv in v.items()} for k, v in osd["state"].items()},
                                                                 "param_groups": osd["param_groups"]})
    fresh, _ = make_agent(hw, dict(case, seed=99), math="fp32", capacity=128)
    ep, steps = fresh.load_checkpoint(p2)
    assert (ep, steps) == (5, 99) and len(fresh.replay_buffer) == 20
    for net in ("policy", "q1", "q2", "q1_target", "q2_target"):
        a, b = net_params(agent, net), net_params(fresh, net)
        for k in a:
            np.testing.assert_array_equal(a[k], b[k])
    assert float(fresh.alpha) == float(agent.alpha) and float(fresh.log_alpha) == float(agent.log_alpha)
    # both agents now take the SAME next step (optimizer state restored bit-for-bit)
    b = batch_of(case, 2)
    l1 = agent.update_from_batch(b, eps=(b["eps_next"], b["eps_cur"]))
    l2 = fresh.update_from_batch(b, eps=(b["eps_next"], b["eps_cur"]))
    assert l1 == l2
    for k, v in net_params(agent, "policy").items():
        np.testing.assert_array_equal(v, net_params(fresh, "policy")[k])
    third, _ = make_agent(hw, dict(case, seed=98), math="fp32")
    third.load(p1)
    np.testing.assert_array_equal(net_params(third, "q2")["fc1.weight"], torch.load(p1, weights_only=False)["q2_state_dict"]["fc1.weight"].cpu().numpy())


def test_seeded_construction_matches_reference_initialisers(hw):
    """torch.manual_seed(s); SAC(...) draws its initial weights through the same torch initialiser calls, in the same
    order, as sac_imp.py:28-36 -> targets start equal to the online critics, policy heads Xavier/orthogonal."""
    hw.use_networks("model2")
    torch.manual_seed(0)
    agent = hw.SAC(17, 6, hidden_dim=64, device="cuda")
    q1, q1t = net_params(agent, "q1"), net_params(agent, "q1_target")
    for k in q1:
        np.testing.assert_array_equal(q1[k], q1t[k])
    w = net_params(agent, "policy")["fc2.weight"]
    np.testing.assert_allclose(w @ w.T, np.eye(64), atol=1e-4)                 # orthogonal init (networks_model2.py:80)
    hw.use_networks("model1")


def test_trainer_loop_call_sequence(hw):
    """The exact call sequence of SACTrainer.train (trainer.py:182-205) against the drop-in classes, with a stub
    environment: select_action -> replay_buffer.push(s, a, r, s2, terminated or truncated) -> `len(buffer) > batch_size`
    gate -> update_parameters(batch_size) -> dict of three python floats appended to the loss history."""
    hw.use_networks("model1")
    obs_dim, act_dim, batch_size = 24, 4, 32
    torch.manual_seed(0); np.random.seed(0); random.seed(0)
    agent = hw.SAC(obs_dim, act_dim, hidden_dim=64, device="cuda", capacity=500, max_batch=batch_size)
    rng = np.random.RandomState(0)
    state = rng.standard_normal(obs_dim)                      # float64 observation, like MuJoCo envs (walk_env.py:30)
    loss_history, total_steps = [], 0
    for total_steps in range(1, 121):
        action = rng.uniform(-0.4, 0.4, act_dim) if total_steps < 20 else agent.select_action(state)       # trainer.py:184-187
        assert action.shape == (act_dim,)
        next_state, reward = rng.standard_normal(obs_dim), float(rng.standard_normal())
        terminated, truncated = bool(rng.uniform() < 0.02), total_steps % 50 == 0
        agent.replay_buffer.push(state, action, reward, next_state, terminated or truncated)                # trainer.py:191-194
        state = next_state
        if len(agent.replay_buffer) > batch_size:                                                           # trainer.py:202
            info = agent.update_parameters(batch_size)                                                      # trainer.py:204
            assert set(info) == {"q1_loss", "q2_loss", "policy_loss"} and all(isinstance(v, float) for v in info.values())
            loss_history.append(info)
    assert len(loss_history) == 120 - batch_size
    assert np.isfinite([v for d in loss_history for v in d.values()]).all()
    assert len(agent.replay_buffer) == 120 and len(agent.replay_buffer.buffer) == 120                      # sac_imp.py:199 reads .buffer
    a_eval = agent.select_action(state, evaluate=True)                                                     # trainer.py:130
    assert np.all(np.abs(a_eval) <= 0.4 + 1e-6)
