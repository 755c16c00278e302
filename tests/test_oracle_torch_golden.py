"""Pins oracle/sac_ref_torch.py (the plain-PyTorch restatement timed as the eager-CUDA baseline) to the live reference's golden
vectors and to the numpy oracle: same seeded inputs, same losses / alpha / updated weights."""
import os

import numpy as np
import pytest

from oracle import sac_oracle_np as O
from oracle import sac_ref_torch as T
from tests.golden import cases

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["tiny_m1", "tiny_m2", "tiny_m1_fixed_alpha", "c1_bipedal_m1"])
def test_torch_restatement_matches_reference(name):
    import torch
    torch.set_num_threads(2)
    case = cases.UPDATE_CASES[name]
    g = np.load(os.path.join(GOLD, f"update_{name}.npz"))
    kw = dict(seed=case["seed"], bias_scale=case.get("bias_scale", 0.0), head_scale=case.get("head_scale", 1.0),
              automatic_entropy_tuning=case.get("auto_entropy", True))
    st = O.make_state(case["obs"], case["act"], case["hidden"], case["n_hidden"], **kw)
    agent = T.TorchSAC(st, "cpu")
    for step in range(case["steps"]):
        b = O.make_batch(case["obs"], case["act"], case["batch"], seed=case["seed"] * 100 + step)
        l, _ = agent.update_from_numpy(b)
        ref = O.update_parameters(st, b)
        np.testing.assert_allclose([l["q1_loss"], l["q2_loss"], l["policy_loss"]], g["losses"][step], rtol=2e-4)
        np.testing.assert_allclose([l[k] for k in ("q1_loss", "q2_loss", "policy_loss")], [ref[k] for k in ("q1_loss", "q2_loss", "policy_loss")], rtol=2e-4)
        np.testing.assert_allclose(float(agent.alpha), g["alphas"][step], rtol=1e-6)
    for net in ("policy", "q1", "q2"):
        for k, v in agent.nets[net].items():
            assert np.abs(v.detach().numpy().ravel()[:64] - g[f"paramhead/{net}/{k}"]).max() < 0.02 * st.lr * case["steps"] + 1e-7
    for net in ("q1", "q2"):
        for k, v in agent.targets[net].items():
            assert np.abs(v.numpy().ravel()[:64] - g[f"paramhead/{net}_target/{k}"]).max() < 0.02 * st.lr * case["steps"] + 1e-7


def test_numpy_per_restatement_matches_reference():
    """NumpyPER.sample == replay_buffer.py:48-68 on the golden case (same np.random seed -> same indices and weights)."""
    case = cases.PER_CASES["small"]
    g = np.load(os.path.join(GOLD, "per_small.npz"))
    pri = cases.per_priorities(case)
    per = T.NumpyPER(pri)
    for call in range(case["calls"]):
        np.random.seed(case["seed"] * 10 + call)
        idx, w = per.sample(case["batch"])
        np.testing.assert_array_equal(idx, g["idx"][call])
        np.testing.assert_allclose(w, g["weights"][call], rtol=1e-6)
