"""CPU restatement (numpy) of the reference SAC learner step.  TEST INFRASTRUCTURE ONLY.

This file is the *oracle* for the B200 path: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package
never does (and fails loudly when its CUDA library is missing).

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the live reference from
``/root/reference`` (possible only in the build container), runs ``SAC.update_parameters`` with
injected minibatch / eps draws and stores the results under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors.

Every function cites the reference lines it follows (paths relative to /root/reference).
All arithmetic is done in ``dtype`` (float32 = the reference's precision, float64 = a truth run
for conditioning checks).  No autograd: the backward pass is written out by hand.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

LOG_SQRT_2PI = math.log(math.sqrt(2.0 * math.pi))  # torch/distributions/normal.py::log_prob
LOG_STD_MIN, LOG_STD_MAX = -20.0, 2.0              # networks_model1.py:74, networks_model2.py:95
SQUASH_EPS = 1e-6                                  # networks_model1.py:96, networks_model2.py:117


# --------------------------------------------------------------------------------------------
# parameter containers
# --------------------------------------------------------------------------------------------
def q_param_names(n_hidden: int) -> List[str]:
    """`Module.parameters()` order of QNetwork: networks_model1.py:15-17 / networks_model2.py:24-27."""
    names = []
    for i in range(1, n_hidden + 2):
        names += [f"fc{i}.weight", f"fc{i}.bias"]
    return names


def policy_param_names(n_hidden: int) -> List[str]:
    """`Module.parameters()` order of GaussianPolicy: networks_model1.py:44-49 / networks_model2.py:56-61."""
    names = []
    for i in range(1, n_hidden + 1):
        names += [f"fc{i}.weight", f"fc{i}.bias"]
    names += ["mean.weight", "mean.bias", "log_std.weight", "log_std.bias"]
    return names


def q_shapes(obs: int, act: int, hidden: int, n_hidden: int) -> Dict[str, Tuple[int, ...]]:
    shp = {}
    fan_in = obs + act
    for i in range(1, n_hidden + 1):
        shp[f"fc{i}.weight"] = (hidden, fan_in)
        shp[f"fc{i}.bias"] = (hidden,)
        fan_in = hidden
    shp[f"fc{n_hidden + 1}.weight"] = (1, hidden)
    shp[f"fc{n_hidden + 1}.bias"] = (1,)
    return shp


def policy_shapes(obs: int, act: int, hidden: int, n_hidden: int) -> Dict[str, Tuple[int, ...]]:
    shp = {}
    fan_in = obs
    for i in range(1, n_hidden + 1):
        shp[f"fc{i}.weight"] = (hidden, fan_in)
        shp[f"fc{i}.bias"] = (hidden,)
        fan_in = hidden
    for head in ("mean", "log_std"):
        shp[f"{head}.weight"] = (act, hidden)
        shp[f"{head}.bias"] = (act,)
    return shp


def xavier_uniform(rng: np.random.RandomState, shape, dtype=np.float32):
    """Same distribution as torch.nn.init.xavier_uniform_ (networks_model1.py:22-25); the stream is
    numpy's (frozen legacy MT19937), so the test inputs regenerate identically on any box."""
    fan_out, fan_in = shape
    bound = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-bound, bound, size=shape).astype(dtype)


def make_net_params(rng, shapes, bias_scale=0.0, dtype=np.float32):
    out = {}
    for name, shp in shapes.items():
        if name.endswith("weight"):
            out[name] = xavier_uniform(rng, shp, dtype)
        else:
            out[name] = (bias_scale * rng.standard_normal(shp)).astype(dtype)
    return out


@dataclass
class AdamState:
    """torch.optim.Adam state for one optimizer (sac_imp.py:39-41,49): per-parameter m, v, shared step."""
    m: Dict[str, np.ndarray]
    v: Dict[str, np.ndarray]
    step: int = 0

    @staticmethod
    def zeros_like(params: Dict[str, np.ndarray]) -> "AdamState":
        return AdamState({k: np.zeros_like(p) for k, p in params.items()},
                         {k: np.zeros_like(p) for k, p in params.items()}, 0)


@dataclass
class SACState:
    """Everything `SAC.__init__` creates (sac_imp.py:9-52) minus the replay buffer."""
    obs: int
    act: int
    hidden: int
    n_hidden: int                      # 2 = networks_model1, 3 = networks_model2
    policy: Dict[str, np.ndarray]
    q1: Dict[str, np.ndarray]
    q2: Dict[str, np.ndarray]
    q1_target: Dict[str, np.ndarray]
    q2_target: Dict[str, np.ndarray]
    policy_opt: AdamState
    q1_opt: AdamState
    q2_opt: AdamState
    log_alpha: np.ndarray              # shape (1,)  sac_imp.py:48
    alpha_opt: AdamState
    alpha: float = 0.2                 # sac_imp.py:23 -- python float until the first update (quirk Q1)
    gamma: float = 0.99
    tau: float = 0.005
    lr: float = 3e-4
    automatic_entropy_tuning: bool = True
    action_scale: float = 0.4          # networks_model1.py:52-55 default bounds (-0.4, 0.4)
    action_bias: float = 0.0
    dtype: type = np.float32

    @property
    def target_entropy(self) -> float:
        return -float(self.act)        # sac_imp.py:46


def make_state(obs, act, hidden, n_hidden, seed=0, dtype=np.float32, bias_scale=0.0, head_scale=1.0, **kw) -> SACState:
    """Deterministic (numpy-seeded) stand-in for `SAC.__init__`; targets start as copies (sac_imp.py:35-36).

    `head_scale` shrinks the mean/log_std head weights.  With N(0,1) observations and Xavier heads the
    pre-tanh action |x_t| exceeds 5 for ~1% of the elements, and there the reference's own fp32 arithmetic
    (`1 - y_t.pow(2)`, networks_model1.py:96) cancels catastrophically: swapping torch's tanh for numpy's
    (<= 1 ulp apart) moves the policy gradient by 1e-2.  Parity cases therefore keep |x_t| < ~4; one
    deliberately saturated case is kept and compared at a correspondingly loose tolerance."""
    rng = np.random.RandomState(seed)
    pol = make_net_params(rng, policy_shapes(obs, act, hidden, n_hidden), bias_scale, dtype)
    for k in ("mean.weight", "log_std.weight"):
        pol[k] = (pol[k] * dtype(head_scale)).astype(dtype)
    q1 = make_net_params(rng, q_shapes(obs, act, hidden, n_hidden), bias_scale, dtype)
    q2 = make_net_params(rng, q_shapes(obs, act, hidden, n_hidden), bias_scale, dtype)
    la = np.zeros(1, dtype)
    return SACState(obs, act, hidden, n_hidden, pol, q1, q2,
                    {k: v.copy() for k, v in q1.items()}, {k: v.copy() for k, v in q2.items()},
                    AdamState.zeros_like(pol), AdamState.zeros_like(q1), AdamState.zeros_like(q2),
                    la, AdamState.zeros_like({"log_alpha": la}), dtype=dtype, **kw)


def make_batch(obs, act, batch, seed=0, dtype=np.float32):
    """Synthetic minibatch + eps draws, BASELINE.md §2 recipe: s,s2~N(0,1), a~U(-.4,.4), r~N(0,1), d~Bern(.01)."""
    rng = np.random.RandomState(1000 + seed)
    s = rng.standard_normal((batch, obs)).astype(dtype)
    a = rng.uniform(-0.4, 0.4, size=(batch, act)).astype(dtype)
    r = rng.standard_normal((batch,)).astype(dtype)
    s2 = rng.standard_normal((batch, obs)).astype(dtype)
    d = (rng.uniform(size=(batch,)) < 0.01).astype(dtype)
    eps_next = rng.standard_normal((batch, act)).astype(dtype)
    eps_cur = rng.standard_normal((batch, act)).astype(dtype)
    return dict(s=s, a=a, r=r, s2=s2, d=d, eps_next=eps_next, eps_cur=eps_cur)


# --------------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------------
class ReluHint:
    """Optional tie-break for ReLU masks (test infrastructure, not part of the reference).

    relu'(z) is discontinuous at z = 0: a pre-activation that lies within rounding of zero is "on" in one
    correct implementation and "off" in another (the reference itself flips such units between BLAS builds),
    and one flipped unit moves a whole row of a weight gradient by O(1/B).  A hint carries the masks another
    implementation used; the oracle adopts them ONLY where its own |z| < band * rms(z) and records every unit
    outside that band where the two disagree (`mismatch`, must stay 0 for parity)."""

    def __init__(self, masks, band=1e-4, band_by_tag=None):
        self.masks, self.band = masks, band        # masks[(tag, layer)] -> bool [B, H]
        self.band_by_tag = band_by_tag or {}       # wider band for passes that run on freshly Adam-stepped weights
        self.adopted = 0                           # units inside the band whose hinted mask differs from z > 0
        self.ambiguous = 0                         # units inside the band
        self.mismatch = 0                          # units OUTSIDE the band where the hint disagrees

    def mask(self, tag, layer, z):
        own = z > 0
        hint = self.masks.get((tag, layer))
        if hint is None:
            return own
        hint = np.asarray(hint, bool).reshape(z.shape)
        inside = np.abs(z) < self.band_by_tag.get(tag, self.band) * max(float(np.sqrt(np.mean(z.astype(np.float64) ** 2))), 1e-30)
        self.ambiguous += int(inside.sum())
        self.adopted += int((inside & (hint != own)).sum())
        self.mismatch += int((~inside & (hint != own)).sum())
        return np.where(inside, hint, own)


def _hidden(P, x, n_hidden, hint, tag):
    """(Linear + ReLU) x n_hidden.  Returns the layer inputs/outputs `acts` and the ReLU masks `masks`
    (masks[i-1] belongs to fc{i}); without a hint masks == (acts > 0), i.e. exactly F.relu and its backward."""
    acts, masks = [x], []
    for i in range(1, n_hidden + 1):
        z = x @ P[f"fc{i}.weight"].T + P[f"fc{i}.bias"]
        if hint is None:
            x = np.maximum(z, 0)
            masks.append(x > 0)
        else:
            m = hint.mask(tag, i, z)
            x = np.where(m, z, 0).astype(z.dtype)
            masks.append(m)
        acts.append(x)
    return x, acts, masks


def q_forward(P, s, a, n_hidden, keep=False, hint=None, tag=None):
    """QNetwork.forward: networks_model1.py:27-33 / networks_model2.py:37-46.  cat(s,a) -> (Linear+ReLU)xn -> Linear(H,1)."""
    x, acts, masks = _hidden(P, np.concatenate([s, a], axis=-1), n_hidden, hint, tag)
    q = x @ P[f"fc{n_hidden + 1}.weight"].T + P[f"fc{n_hidden + 1}.bias"]      # [B,1]
    return (q, (acts, masks)) if keep else q


def policy_forward(P, s, n_hidden, keep=False, hint=None, tag=None):
    """GaussianPolicy.forward: networks_model1.py:65-76 / networks_model2.py:85-97."""
    x, acts, masks = _hidden(P, s, n_hidden, hint, tag)
    mean = x @ P["mean.weight"].T + P["mean.bias"]
    ls_raw = x @ P["log_std.weight"].T + P["log_std.bias"]
    log_std = np.clip(ls_raw, LOG_STD_MIN, LOG_STD_MAX)
    if keep:
        return mean, log_std, ls_raw, (acts, masks)
    return mean, log_std


def policy_sample(P, s, eps, n_hidden, scale, bias, keep=False, hint=None, tag=None):
    """GaussianPolicy.sample: networks_model1.py:78-99 == networks_model2.py:99-120, with the N(0,1) draw
    of `Normal.rsample` (x_t = mean + eps*std) supplied by the caller."""
    dt = s.dtype.type
    mean, log_std, ls_raw, acts = policy_forward(P, s, n_hidden, keep=True, hint=hint, tag=tag)
    std = np.exp(log_std)
    x_t = mean + eps * std
    y_t = np.tanh(x_t)
    action = y_t * dt(scale) + dt(bias)
    var = std * std
    log_scale = np.log(std)
    log_prob = -((x_t - mean) ** 2) / (dt(2) * var) - log_scale - dt(LOG_SQRT_2PI)
    log_prob = log_prob - np.log(dt(scale) * (dt(1) - y_t * y_t) + dt(SQUASH_EPS))
    log_prob = log_prob.sum(axis=-1, keepdims=True)
    if keep:
        return action, log_prob, dict(mean=mean, log_std=log_std, ls_raw=ls_raw, std=std, y=y_t, acts=acts)
    return action, log_prob


def select_action(st: SACState, state, evaluate=False, eps=None):
    """SAC.select_action: sac_imp.py:54-72 (B=1)."""
    s = np.asarray(state, st.dtype)[None, :]
    if evaluate:
        mean, _ = policy_forward(st.policy, s, st.n_hidden)
        return (np.tanh(mean) * st.dtype(st.action_scale) + st.dtype(st.action_bias))[0]
    a, _ = policy_sample(st.policy, s, np.asarray(eps, st.dtype)[None, :], st.n_hidden,
                         st.action_scale, st.action_bias)
    return a[0]


# --------------------------------------------------------------------------------------------
# backward pieces (hand-written; validated against the reference's autograd in make_golden.py)
# --------------------------------------------------------------------------------------------
def q_backward(P, acts_masks, dq, n_hidden, need_dw=True):
    """Backward of q_forward given dL/dq [B,1].  Returns (grads or None, dL/dx of the cat(s,a) input)."""
    acts, masks = acts_masks
    grads = {}
    L = n_hidden + 1
    g = dq
    for i in range(L, 0, -1):
        x_in = acts[i - 1]
        W = P[f"fc{i}.weight"]
        if need_dw:
            grads[f"fc{i}.weight"] = g.T @ x_in
            grads[f"fc{i}.bias"] = g.sum(axis=0)
        g = g @ W
        if i > 1:
            g = g * masks[i - 2]          # relu'(z_{i-1}) == (x_in > 0) unless a ReluHint overrode it
    return (grads if need_dw else None), g


def policy_backward(P, acts_masks, g_mean, g_ls, n_hidden):
    acts, masks = acts_masks
    grads = {
        "mean.weight": g_mean.T @ acts[-1], "mean.bias": g_mean.sum(axis=0),
        "log_std.weight": g_ls.T @ acts[-1], "log_std.bias": g_ls.sum(axis=0),
    }
    g = (g_mean @ P["mean.weight"] + g_ls @ P["log_std.weight"]) * masks[-1]
    for i in range(n_hidden, 0, -1):
        x_in = acts[i - 1]
        grads[f"fc{i}.weight"] = g.T @ x_in
        grads[f"fc{i}.bias"] = g.sum(axis=0)
        if i > 1:
            g = (g @ P[f"fc{i}.weight"]) * masks[i - 2]
    return grads


def adam_step(params, grads, opt: AdamState, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch/optim/adam.py::_single_tensor_adam, non-capturable branch (defaults of sac_imp.py:39-41):
    m.lerp_(g,1-b1); v = b2 v + (1-b2) g^2; denom = sqrt(v)/sqrt(1-b2^t) + eps; p -= lr/(1-b1^t) * m/denom."""
    opt.step += 1
    t = opt.step
    bc1 = 1.0 - beta1 ** t
    bc2 = 1.0 - beta2 ** t
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for k, p in params.items():
        dt = p.dtype.type
        g = grads[k].reshape(p.shape).astype(p.dtype)
        m, v = opt.m[k], opt.v[k]
        m += dt(1.0 - beta1) * (g - m)
        v *= dt(beta2)
        v += dt(1.0 - beta2) * g * g
        denom = np.sqrt(v) / dt(bc2_sqrt) + dt(eps)
        p -= dt(step_size) * (m / denom)


def polyak(target, online, tau):
    """SAC._soft_update_target_networks: sac_imp.py:146-152."""
    for k in target:
        dt = target[k].dtype.type
        target[k][...] = target[k] * dt(1.0 - tau) + online[k] * dt(tau)


# --------------------------------------------------------------------------------------------
# the learner step
# --------------------------------------------------------------------------------------------
def update_parameters(st: SACState, batch, per_weights: Optional[np.ndarray] = None, return_aux=False,
                      relu_hint: Optional[ReluHint] = None):
    """SAC.update_parameters: sac_imp.py:74-144, same order of operations.

    `batch` = dict(s,a,r,s2,d,eps_next,eps_cur); r,d are [B] and are unsqueezed as at sac_imp.py:83,85.
    `per_weights` (extension H10, not in the reference): IS weights w[B]; critic loss becomes mean(w (q-y)^2).
    `relu_hint` (test infrastructure): masks of another implementation for the three passes that are differentiated,
    tags 'q1'/'q2' (critics on (s,a)), 'q1a'/'q2a' (updated critics on (s,a_new)), 'policy' (policy on s); see ReluHint.
    Returns {'q1_loss','q2_loss','policy_loss'} (+ aux dict with grads/targets/td when return_aux)."""
    dt = st.dtype
    s, a, s2 = batch["s"].astype(dt), batch["a"].astype(dt), batch["s2"].astype(dt)
    r = batch["r"].astype(dt)[:, None]
    d = batch["d"].astype(dt)[:, None]
    B = s.shape[0]
    nh = st.n_hidden
    alpha = dt(st.alpha)

    # -- target (no grad)  sac_imp.py:87-98
    a2, logp2 = policy_sample(st.policy, s2, batch["eps_next"].astype(dt), nh, st.action_scale, st.action_bias)
    q1n = q_forward(st.q1_target, s2, a2, nh)
    q2n = q_forward(st.q2_target, s2, a2, nh)
    qn = np.minimum(q1n, q2n)
    value_target = qn - alpha * logp2
    y = r + (dt(1) - d) * dt(st.gamma) * value_target

    # -- critics  sac_imp.py:101-113
    aux = {}
    losses = {}
    w = None if per_weights is None else per_weights.astype(dt)[:, None]
    td = []
    for name, P, opt in (("q1", st.q1, st.q1_opt), ("q2", st.q2, st.q2_opt)):
        qp, acts = q_forward(P, s, a, nh, keep=True, hint=relu_hint, tag=name)
        diff = qp - y
        td.append(diff[:, 0].copy())
        if w is None:
            losses[f"{name}_loss"] = float(np.mean(diff * diff))         # F.mse_loss, reduction=mean over [B,1]
            dq = dt(2) * diff / dt(B)
        else:
            losses[f"{name}_loss"] = float(np.mean(w * diff * diff))
            dq = dt(2) * w * diff / dt(B)
        grads, _ = q_backward(P, acts, dq, nh)
        aux[f"{name}_grads"] = grads
        adam_step(P, grads, opt, st.lr)

    # -- actor  sac_imp.py:116-125 (uses the UPDATED q1,q2; alpha and Q weights are constants here, quirk Q2)
    an, logp, pk = policy_sample(st.policy, s, batch["eps_cur"].astype(dt), nh, st.action_scale, st.action_bias, keep=True,
                                 hint=relu_hint, tag="policy")
    q1p, acts1 = q_forward(st.q1, s, an, nh, keep=True, hint=relu_hint, tag="q1a")
    q2p, acts2 = q_forward(st.q2, s, an, nh, keep=True, hint=relu_hint, tag="q2a")
    qmin = np.minimum(q1p, q2p)
    losses["policy_loss"] = float(np.mean(alpha * logp - qmin))
    # d(-mean(min))/dq_k : routed to the smaller Q, exact ties split 1/2-1/2 (torch.minimum backward)
    sel1 = (q1p < q2p).astype(dt) + dt(0.5) * (q1p == q2p).astype(dt)
    dq1 = -sel1 / dt(B)
    dq2 = -(dt(1) - sel1) / dt(B)
    _, dx1 = q_backward(st.q1, acts1, dq1, nh, need_dw=False)
    _, dx2 = q_backward(st.q2, acts2, dq2, nh, need_dw=False)
    dLda = (dx1 + dx2)[:, st.obs:]
    y_t, std = pk["y"], pk["std"]
    sc = dt(st.action_scale)
    one_m_y2 = dt(1) - y_t * y_t
    g_u = dLda * sc * one_m_y2 + (alpha / dt(B)) * (dt(2) * sc * y_t * one_m_y2) / (sc * one_m_y2 + dt(SQUASH_EPS))
    g_mean = g_u
    in_range = ((pk["ls_raw"] >= LOG_STD_MIN) & (pk["ls_raw"] <= LOG_STD_MAX)).astype(dt)
    g_ls = (g_u * std * batch["eps_cur"].astype(dt) - alpha / dt(B)) * in_range
    pgrads = policy_backward(st.policy, pk["acts"], g_mean, g_ls, nh)
    aux["policy_grads"] = pgrads
    adam_step(st.policy, pgrads, st.policy_opt, st.lr)

    # -- temperature  sac_imp.py:128-135
    if st.automatic_entropy_tuning:
        g_la = -np.mean(logp + dt(st.target_entropy), dtype=dt).reshape(1)
        aux["log_alpha_grad"] = g_la.copy()
        aux["alpha_loss"] = float(-(st.log_alpha[0] * np.mean(logp + dt(st.target_entropy))))
        adam_step({"log_alpha": st.log_alpha}, {"log_alpha": g_la}, st.alpha_opt, st.lr)
        st.alpha = float(np.exp(st.log_alpha[0]))

    # -- Polyak  sac_imp.py:138
    polyak(st.q1_target, st.q1, st.tau)
    polyak(st.q2_target, st.q2, st.tau)

    if return_aux:
        aux.update(y=y[:, 0], td1=td[0], td2=td[1], logp=logp[:, 0], logp_next=logp2[:, 0],
                   action_new=an, action_next=a2)
        return losses, aux
    return losses


