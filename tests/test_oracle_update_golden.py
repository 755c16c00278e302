"""Pins oracle/sac_oracle_np.py to the live reference: golden vectors made by tests/golden/make_golden.py
(reference sac_imp.SAC.update_parameters, sac_imp.py:74-144, run with injected minibatch / eps)."""
import os

import numpy as np
import pytest

from oracle import sac_oracle_np as O
from tests.golden import cases

GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 2e-4   # fp32 reassociation between numpy/OpenBLAS here and torch/MKL autograd in the reference


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def replay_oracle(case, dtype=np.float32):
    st = O.make_state(case["obs"], case["act"], case["hidden"], case["n_hidden"], seed=case["seed"],
                      bias_scale=case.get("bias_scale", 0.0), head_scale=case.get("head_scale", 1.0), dtype=dtype,
                      automatic_entropy_tuning=case.get("auto_entropy", True))
    losses, alphas, aux0 = [], [], None
    for step in range(case["steps"]):
        b = O.make_batch(case["obs"], case["act"], case["batch"], seed=case["seed"] * 100 + step)
        l, aux = O.update_parameters(st, b, return_aux=True)
        if step == 0:
            aux0 = aux
        losses.append([l["q1_loss"], l["q2_loss"], l["policy_loss"]])
        alphas.append(st.alpha)
    return st, np.array(losses), np.array(alphas), aux0


@pytest.mark.parametrize("name", list(cases.UPDATE_CASES))
def test_update_matches_reference(name):
    case = cases.UPDATE_CASES[name]
    g = np.load(os.path.join(GOLD, f"update_{name}.npz"))
    st, losses, alphas, aux = replay_oracle(case)
    if case.get("loose"):
        # saturated tanh: the reference's own fp32 result moves by ~1e-2 with a 1-ulp change of tanh (see make_state)
        np.testing.assert_allclose(losses, g["losses"], rtol=2e-2)
        for nm, gr in aux["policy_grads"].items():
            ref = g[f"gradsum/policy/{nm}"]
            assert abs(np.sqrt((gr.astype(np.float64) ** 2).sum()) - ref[1]) <= 5e-2 * ref[1], nm
        return
    np.testing.assert_allclose(losses, g["losses"], rtol=RTOL)
    np.testing.assert_allclose(alphas, g["alphas"], rtol=1e-6)
    for net in ("q1", "q2", "policy"):
        for nm, gr in aux[f"{net}_grads"].items():
            ref = g[f"gradsum/{net}/{nm}"]
            assert abs(np.sqrt((gr.astype(np.float64) ** 2).sum()) - ref[1]) <= RTOL * ref[1] + 1e-12, (net, nm)
            if case["full"]:
                assert relerr(gr.reshape(g[f"grad/{net}/{nm}"].shape), g[f"grad/{net}/{nm}"]) < RTOL, (net, nm)
    if case.get("auto_entropy", True):
        np.testing.assert_allclose(aux["log_alpha_grad"], g["grad/log_alpha"], rtol=RTOL)
        np.testing.assert_allclose(st.log_alpha, g["log_alpha"], rtol=1e-5, atol=1e-9)
    for net in ("policy", "q1", "q2", "q1_target", "q2_target"):
        for nm, p in getattr(st, net).items():
            # Adam's first steps move every weight by ~lr regardless of |g|: compare absolutely, in units of lr
            head = g[f"paramhead/{net}/{nm}"]
            assert np.abs(p.ravel()[:64] - head).max() < 0.02 * st.lr * case["steps"] + 1e-7, (net, nm)
            if case["full"]:
                assert np.abs(p - g[f"param/{net}/{nm}"]).max() < 0.02 * st.lr * case["steps"] + 1e-7, (net, nm)
    if case["full"]:
        for net, opt in (("policy", st.policy_opt), ("q1", st.q1_opt), ("q2", st.q2_opt)):
            for nm in opt.m:
                assert relerr(opt.m[nm], g[f"adam_m/{net}/{nm}"]) < RTOL
                assert relerr(opt.v[nm], g[f"adam_v/{net}/{nm}"]) < 2 * RTOL


@pytest.mark.parametrize("name", ["tiny_m1", "tiny_m2", "c2_humanoid_m2"])
def test_select_action_matches_reference(name):
    case = cases.UPDATE_CASES[name]
    g = np.load(os.path.join(GOLD, f"update_{name}.npz"))
    st, *_ = replay_oracle(case)
    obs_vec = np.random.RandomState(77 + case["seed"]).standard_normal(case["obs"]).astype(np.float32)
    eps_vec = np.random.RandomState(78 + case["seed"]).standard_normal((1, case["act"])).astype(np.float32)
    np.testing.assert_allclose(O.select_action(st, obs_vec, evaluate=True), g["select/eval"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(O.select_action(st, obs_vec, eps=eps_vec[0]), g["select/sample"], rtol=1e-3, atol=1e-5)


def test_float64_oracle_agrees_with_float32():
    case = cases.UPDATE_CASES["tiny_m2"]
    _, l32, _, _ = replay_oracle(case, np.float32)
    _, l64, _, _ = replay_oracle(case, np.float64)
    np.testing.assert_allclose(l32, l64, rtol=1e-3)
