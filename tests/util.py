"""Shared helpers of the GPU parity tests: load an oracle state into the product SAC, compare tensors."""
import numpy as np
import torch

from oracle import sac_oracle_np as O


def relerr(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_agent(hw, case, math="fp32", launch="staged", **kw):
    """Product SAC holding exactly the numpy-seeded weights of oracle make_state(case)."""
    hw.use_networks("model1" if case["n_hidden"] == 2 else "model2")
    st = O.make_state(case["obs"], case["act"], case["hidden"], case["n_hidden"], seed=case["seed"],
                      bias_scale=case.get("bias_scale", 0.0), head_scale=case.get("head_scale", 1.0),
                      automatic_entropy_tuning=case.get("auto_entropy", True))
    agent = hw.SAC(case["obs"], case["act"], hidden_dim=case["hidden"], device="cuda", math=math, launch=launch,
                   automatic_entropy_tuning=case.get("auto_entropy", True), max_batch=max(case["batch"], 16),
                   capacity=kw.pop("capacity", 4096), seed=1234, **kw)
    for net in ("policy", "q1", "q2", "q1_target", "q2_target"):
        getattr(agent, net).load_state_dict({k: torch.from_numpy(v.copy()) for k, v in getattr(st, net).items()})
    return agent, st


def relu_hint(agent, case, band=1e-4, band_post_adam=1e-2):
    """ReLU masks the device used in the last update (tests/oracle ReluHint): the three differentiated passes.

    band: |z| < 1e-4 rms(z) may take either sign (bf16-pair operand storage, 2^-16 per element).  The actor phase
    evaluates the critics right AFTER their Adam step (sac_imp.py:109,113 -> :117-118): a weight whose gradient is
    within rounding of zero moves by +lr or -lr (early Adam steps are sign-like), shifting a pre-activation by up to
    2*lr*|x|; those two passes get the wider band."""
    import ctypes
    from humanoid_walking_with_sac_b200 import _native as N
    B, H, nh = case["batch"], case["hidden"], case["n_hidden"]
    masks = {}
    buf = np.empty((B, H), np.float32)
    for layer in range(nh):
        for tag, group, k in (("policy", 0, 0), ("q1", 1, 0), ("q2", 1, 1), ("q1a", 2, 0), ("q2a", 2, 1)):
            N.check(N.lib().sacb_debug_read_activation(agent._h, 0, group, k, layer, B, N.ptr(buf)))
            masks[(tag, layer + 1)] = buf > 0
    return O.ReluHint(masks, band, {"q1a": band_post_adam, "q2a": band_post_adam})


def batch_of(case, step):
    return O.make_batch(case["obs"], case["act"], case["batch"], seed=case["seed"] * 100 + step)


def net_params(agent, net):
    return {k: v.detach().cpu().numpy() for k, v in getattr(agent, net).state_dict().items()}


def grad_close(got, ref, tol, max_flips=2):
    """Gradient parity that tolerates a couple of ReLU mask flips.

    A hidden unit whose pre-activation lies within rounding (~1e-6) of zero for one sample is on for one
    summation order and off for another; the reference itself flips such units when its BLAS changes.  One flip
    moves one row of that layer's weight gradient (and one bias entry) by O(1/B).  So: every row of the tensor must
    match within `tol` (relative to the RMS row norm) except at most `max_flips` rows, and the whole tensor within 5e-2."""
    got = np.asarray(got, np.float64).reshape(np.asarray(ref).shape)
    ref = np.asarray(ref, np.float64)
    g2, r2 = (got.reshape(got.shape[0], -1), ref.reshape(ref.shape[0], -1)) if ref.ndim == 2 and ref.shape[0] > 1 else (got.reshape(-1, 1), ref.reshape(-1, 1))
    row_err = np.linalg.norm(g2 - r2, axis=1)
    scale = max(np.sqrt(np.mean(np.sum(r2 * r2, axis=1))), 1e-30)
    bad = int(np.sum(row_err > tol * scale * max(1.0, np.sqrt(r2.shape[0]) / 4)))
    overall = relerr(got, ref)
    ok = overall < tol or (bad <= max_flips and overall < 5e-2)
    return ok, overall, bad


def state_from_checkpoint(path, case, alpha_from_file=True):
    """Oracle SACState holding the networks of a reference-format `save()` / `save_checkpoint()` file (sac_imp.py:154-233)."""
    ck = torch.load(path, map_location="cpu", weights_only=False)
    st = O.make_state(case["obs"], case["act"], case["hidden"], case["n_hidden"], seed=0)
    for net in ("policy", "q1", "q2", "q1_target", "q2_target"):
        setattr(st, net, {k: v.detach().cpu().numpy().astype(np.float32).copy() for k, v in ck[f"{net}_state_dict"].items()})
    if alpha_from_file:
        a = ck["alpha"]
        st.alpha = float(a.detach().reshape(-1)[0]) if torch.is_tensor(a) else float(a)
    for net, opt in (("policy", "policy_opt"), ("q1", "q1_opt"), ("q2", "q2_opt")):
        sd = ck.get(f"{net}_optimizer_state_dict")
        if sd and sd["state"]:
            names = list(getattr(st, net).keys())
            o = getattr(st, opt)
            for i, nm in enumerate(names):
                o.m[nm] = sd["state"][i]["exp_avg"].cpu().numpy().copy()
                o.v[nm] = sd["state"][i]["exp_avg_sq"].cpu().numpy().copy()
            o.step = int(float(sd["state"][0]["step"]))
    if "log_alpha" in ck:
        st.log_alpha = ck["log_alpha"].detach().cpu().numpy().astype(np.float32).reshape(1).copy()
        sd = ck.get("alpha_optimizer_state_dict")
        if sd and sd["state"]:
            st.alpha_opt.m["log_alpha"] = sd["state"][0]["exp_avg"].cpu().numpy().reshape(1).copy()
            st.alpha_opt.v["log_alpha"] = sd["state"][0]["exp_avg_sq"].cpu().numpy().reshape(1).copy()
            st.alpha_opt.step = int(float(sd["state"][0]["step"]))
    return st, ck
