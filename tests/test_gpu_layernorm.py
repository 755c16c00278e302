"""Opt-in LayerNorm variant (`SAC(..., layer_norm=True)`: Linear -> LayerNorm -> ReLU in every hidden layer of all five networks).

NO REFERENCE PARITY: the reference has no LayerNorm (networks_model2.py:86 is only a comment).  The oracle
here is test-local: the plain-PyTorch restatement of the learner step (oracle/sac_ref_torch.py, pinned to the reference's golden
vectors WITHOUT LayerNorm) with torch.nn.functional.layer_norm between each hidden Linear and its ReLU, differentiated by autograd.
Default off: every other test runs the reference architecture."""
import numpy as np
import pytest
import torch

from oracle import sac_oracle_np as O
from oracle import sac_ref_torch as T

pytestmark = pytest.mark.gpu

CASES = {
    "tiny_m2": dict(obs=11, act=3, hidden=64, n_hidden=3, batch=32, steps=3, seed=5),
    "ragged_m1": dict(obs=24, act=4, hidden=72, n_hidden=2, batch=37, steps=2, seed=6),
    "c2_like_m2": dict(obs=348, act=17, hidden=512, n_hidden=3, batch=256, steps=2, seed=7),
    "large_batch_m2": dict(obs=348, act=17, hidden=512, n_hidden=3, batch=2048, steps=1, seed=8),      # throughput form (stream kernel + split stages)
}


@pytest.fixture(scope="module")
def hw():
    import humanoid_walking_with_sac_b200 as hw
    return hw


def _agent_and_oracle(hw, case, launch="staged"):
    hw.use_networks("model1" if case["n_hidden"] == 2 else "model2")
    torch.manual_seed(case["seed"])
    agent = hw.SAC(case["obs"], case["act"], hidden_dim=case["hidden"], device="cuda", max_batch=max(case["batch"], 16), capacity=1024,
                   seed=99, layer_norm=True, launch=launch)
    rng = np.random.RandomState(case["seed"])
    st = O.make_state(case["obs"], case["act"], case["hidden"], case["n_hidden"], seed=case["seed"], bias_scale=0.05, head_scale=0.5)
    for net in ("policy", "q1", "q2", "q1_target", "q2_target"):
        sd = dict(getattr(st, net))
        src = net.replace("_target", "")
        for i in range(1, case["n_hidden"] + 1):      # non-trivial gamma / beta; a target starts as a copy of its critic (sac_imp.py:37-38)
            if net.endswith("_target"):
                sd[f"ln{i}.weight"], sd[f"ln{i}.bias"] = getattr(st, src)[f"ln{i}.weight"].copy(), getattr(st, src)[f"ln{i}.bias"].copy()
            else:
                sd[f"ln{i}.weight"] = (1.0 + 0.2 * rng.standard_normal(case["hidden"])).astype(np.float32)
                sd[f"ln{i}.bias"] = (0.1 * rng.standard_normal(case["hidden"])).astype(np.float32)
        keys = list(getattr(agent, net).state_dict().keys())
        sd = {k: sd[k] for k in keys}                 # module order: fc1.*, ln1.*, fc2.*, ...
        setattr(st, net, sd)
        getattr(agent, net).load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd.items()})
    return agent, st, T.TorchSAC(st, "cpu")


@pytest.mark.parametrize("name", list(CASES))
def test_layernorm_update_matches_torch_autograd(hw, name):
    case = CASES[name]
    agent, st, ref = _agent_and_oracle(hw, case)
    assert [k for k in agent.q1.state_dict()][:4] == ["fc1.weight", "fc1.bias", "ln1.weight", "ln1.bias"]
    lr = st.lr
    for step in range(case["steps"]):
        b = O.make_batch(case["obs"], case["act"], case["batch"], seed=case["seed"] * 100 + step)
        got = agent.update_from_batch(b, eps=(b["eps_next"], b["eps_cur"]))
        want, _ = ref.update_from_numpy(b)
        for k in ("q1_loss", "q2_loss", "policy_loss"):
            assert abs(got[k] - want[k]) <= 5e-4 * abs(want[k]) + 1e-6, (step, k, got[k], want[k])
    a = agent.alpha
    assert abs(float(a.item() if hasattr(a, "item") else a) - float(ref.alpha)) <= 1e-5 * abs(float(ref.alpha))
    # weights (gamma / beta included), targets: within a small fraction of an Adam step for all but a few per cent of the elements (a unit
    # whose pre-activation is within rounding of zero takes either ReLU branch -- tests/util.py::grad_close -- and Adam's first steps are
    # sign-like, so such a flip moves a whole weight row by a good part of a step; no ReLU hint is handed to this oracle)
    budget = 0.05 * lr * case["steps"] + 1e-7
    for net, params in (("policy", ref.nets["policy"]), ("q1", ref.nets["q1"]), ("q2", ref.nets["q2"]), ("q1_target", ref.targets["q1"]), ("q2_target", ref.targets["q2"])):
        mine = {k: v.detach().cpu().numpy() for k, v in getattr(agent, net).state_dict().items()}
        assert set(mine) == set(params)
        for k, v in params.items():
            d = np.abs(mine[k] - v.detach().numpy())
            assert d.max() <= 2.1 * lr * case["steps"] + 1e-6 and np.mean(d > budget) < 0.05, (net, k, float(d.max()), float(np.mean(d > budget)))
    moved = agent.q1.state_dict()["ln1.weight"].cpu().numpy() - st.q1["ln1.weight"]
    assert np.abs(moved).max() > 0.5 * lr      # gamma really is trained


def test_layernorm_acting_and_forward_match_torch(hw):
    case = CASES["tiny_m2"]
    agent, st, ref = _agent_and_oracle(hw, case)
    rng = np.random.RandomState(3)
    s = rng.standard_normal((5, case["obs"])).astype(np.float32)
    a = rng.uniform(-0.4, 0.4, (5, case["act"])).astype(np.float32)
    with torch.no_grad():
        q_ref = ref.q(ref.nets["q1"], torch.from_numpy(s), torch.from_numpy(a)).numpy()
        h = ref._trunk(ref.nets["policy"], torch.from_numpy(s))
        mean_ref = torch.nn.functional.linear(h, ref.nets["policy"]["mean.weight"], ref.nets["policy"]["mean.bias"]).numpy()
    np.testing.assert_allclose(agent.q1(torch.from_numpy(s), torch.from_numpy(a)).numpy(), q_ref, rtol=2e-4, atol=2e-5)
    mean, _ = agent.policy(torch.from_numpy(s))
    np.testing.assert_allclose(mean.numpy(), mean_ref, rtol=2e-4, atol=2e-5)
    act = agent.select_action(s[0], evaluate=True)
    np.testing.assert_allclose(act, np.tanh(mean_ref[0]) * st.action_scale + st.action_bias, rtol=2e-4, atol=2e-5)


def test_layernorm_persistent_equals_staged_and_checkpoint_round_trip(hw, tmp_path):
    case = CASES["tiny_m2"]
    a1, _, _ = _agent_and_oracle(hw, case, launch="staged")
    a2, _, _ = _agent_and_oracle(hw, case, launch="persistent")
    for step in range(2):
        b = O.make_batch(case["obs"], case["act"], case["batch"], seed=900 + step)
        l1 = a1.update_from_batch(b, eps=(b["eps_next"], b["eps_cur"]))
        l2 = a2.update_from_batch(b, eps=(b["eps_next"], b["eps_cur"]))
        assert l1 == l2
    for net in ("policy", "q1", "q2_target"):
        for k, v in getattr(a1, net).state_dict().items():
            assert torch.equal(v, getattr(a2, net).state_dict()[k]), (net, k)
    path = str(tmp_path / "ln.pt")
    a1.save_checkpoint(path, 1, 2)
    hw.use_networks("model2")
    fresh = hw.SAC(case["obs"], case["act"], hidden_dim=case["hidden"], device="cuda", max_batch=max(case["batch"], 16), capacity=1024, seed=99, layer_norm=True)
    fresh.load_checkpoint(path)
    b = O.make_batch(case["obs"], case["act"], case["batch"], seed=950)
    assert fresh.update_from_batch(b, eps=(b["eps_next"], b["eps_cur"])) == a1.update_from_batch(b, eps=(b["eps_next"], b["eps_cur"]))


def test_layernorm_is_refused_where_it_is_not_validated(hw):
    """population handles and the data-parallel entry points refuse the LayerNorm variant instead of running unvalidated programs"""
    import ctypes
    N = hw._native
    lib = N.lib()
    cfg = N.default_config()
    cfg.obs_dim, cfg.act_dim, cfg.hidden_dim, cfg.n_hidden, cfg.n_agents, cfg.layer_norm, cfg.capacity = 8, 2, 64, 2, 2, 1, 1024
    h = ctypes.c_void_p()
    assert lib.sacb_create(ctypes.byref(cfg), ctypes.byref(h)) == N.ERR_ARG
    agent = hw.SAC(8, 2, 64, layer_norm=True, capacity=1024, seed=1)
    with pytest.raises(ValueError):
        hw.distributed.DataParallelSAC(agent)
