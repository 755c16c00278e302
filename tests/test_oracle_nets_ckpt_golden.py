"""Pins the oracle's network callables and its handling of reference-written checkpoints to the live reference
(fixtures: tests/golden/nets_*.npz, ckpt_ref_*.pt, shipped_*_best_model.pt + *_expected.npz, all made by make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import sac_oracle_np as O
from tests.golden import cases
from tests.util import state_from_checkpoint

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SHIPPED = {"bipedal": dict(obs=24, act=4, hidden=256, n_hidden=2, batch=256, seed=31),
           "humanoid376": dict(obs=376, act=17, hidden=256, n_hidden=2, batch=256, seed=32)}


@pytest.mark.parametrize("name", cases.NETS_CASES)
def test_network_callables_match_reference(name):
    """QNetwork.forward / GaussianPolicy.forward / .sample (networks_model1.py:27-33, :65-99; networks_model2.py:37-46, :85-120)."""
    case = cases.UPDATE_CASES[name]
    g = np.load(os.path.join(GOLD, f"nets_{name}.npz"))
    st = O.make_state(case["obs"], case["act"], case["hidden"], case["n_hidden"], seed=case["seed"],
                      bias_scale=case.get("bias_scale", 0.0), head_scale=case.get("head_scale", 1.0))
    inp = cases.nets_inputs(case, 32)
    np.testing.assert_allclose(O.q_forward(st.q1, inp["s"], inp["a"], st.n_hidden), g["q1"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(O.q_forward(st.q2_target, inp["s"], inp["a"], st.n_hidden), g["q2_target"], rtol=2e-5, atol=2e-6)
    mean, log_std = O.policy_forward(st.policy, inp["s"], st.n_hidden)
    np.testing.assert_allclose(mean, g["mean"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(log_std, g["log_std"], rtol=2e-5, atol=2e-6)
    action, logp = O.policy_sample(st.policy, inp["s"], inp["eps"], st.n_hidden, st.action_scale, st.action_bias)
    np.testing.assert_allclose(action, g["action"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(logp, g["log_prob"], rtol=1e-4, atol=1e-4)


def test_oracle_resumes_a_reference_checkpoint():
    """save_checkpoint() file written by the reference (sac_imp.py:177-201) -> oracle state -> the reference's own next step."""
    case = cases.UPDATE_CASES["tiny_m1"]
    exp = np.load(os.path.join(GOLD, "ckpt_ref_expected.npz"))
    st, ck = state_from_checkpoint(os.path.join(GOLD, "ckpt_ref_checkpoint.pt"), case)
    assert (ck["episode"], ck["total_steps"]) == (7, 123) and len(ck["replay_buffer"]) == 20
    b = O.make_batch(case["obs"], case["act"], case["batch"], seed=case["seed"] * 100 + 2)
    l = O.update_parameters(st, b)
    np.testing.assert_allclose([l["q1_loss"], l["q2_loss"], l["policy_loss"]], exp["next_losses"][0], rtol=2e-4)
    np.testing.assert_allclose(st.alpha, exp["next_alpha"], rtol=1e-6)


@pytest.mark.parametrize("tag", list(SHIPPED))
def test_oracle_on_shipped_checkpoints(tag):
    """The reference's shipped results/*/best_model.pt (obs 24 and obs 376): act + one seeded update, oracle vs live reference."""
    case = SHIPPED[tag]
    exp = np.load(os.path.join(GOLD, f"shipped_{tag}_expected.npz"))
    st, _ = state_from_checkpoint(os.path.join(GOLD, f"shipped_{tag}_best_model.pt"), case)
    np.testing.assert_allclose(st.alpha, exp["alpha_loaded"], rtol=1e-7)
    obs_mat = np.random.RandomState(77 + case["seed"]).standard_normal((8, case["obs"])).astype(np.float32)
    got = np.stack([O.select_action(st, o, evaluate=True) for o in obs_mat])
    np.testing.assert_allclose(got, exp["select_eval"], rtol=1e-4, atol=1e-5)
    inp = cases.nets_inputs(case, 16)
    np.testing.assert_allclose(O.q_forward(st.q1, inp["s"], inp["a"], 2), exp["q1"], rtol=1e-4, atol=1e-4)
    b = O.make_batch(case["obs"], case["act"], case["batch"], seed=case["seed"] * 100)
    l = O.update_parameters(st, b)
    np.testing.assert_allclose([l["q1_loss"], l["q2_loss"]], exp["losses"][0][:2], rtol=5e-4)
    np.testing.assert_allclose(l["policy_loss"], exp["losses"][0][2], rtol=5e-3)      # trained heads saturate tanh on N(0,1) observations
    np.testing.assert_allclose(st.alpha, exp["alpha_after"], rtol=1e-5)
