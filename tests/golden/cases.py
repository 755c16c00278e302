"""Golden-case definitions shared by make_golden.py (reference side) and the tests (product side).

Inputs regenerate from seeds (numpy legacy RandomState: stream frozen by NEP 19), outputs are
stored in the .npz next to this file.  Shapes follow BASELINE.json:configs (C1..C3) plus two tiny
full-state cases that keep every tensor.
"""
import numpy as np

UPDATE_CASES = {
    # tiny, every tensor kept (weights, Adam m/v, grads) -- model1 = 2 hidden, model2 = 3 hidden
    "tiny_m1": dict(obs=11, act=3, hidden=32, n_hidden=2, batch=16, steps=3, seed=1, bias_scale=0.05, full=True),
    "tiny_m2": dict(obs=13, act=5, hidden=48, n_hidden=3, batch=24, steps=3, seed=2, bias_scale=0.05, full=True),
    "tiny_m1_fixed_alpha": dict(obs=7, act=2, hidden=16, n_hidden=2, batch=8, steps=3, seed=3, bias_scale=0.05,
                                full=True, auto_entropy=False),
    # BASELINE.json configs: summaries only (sum / L2 / max per tensor + first 64 values)
    # head_scale: keeps the pre-tanh action out of the saturated region where the reference's fp32 is chaotic
    "c1_bipedal_m1": dict(obs=24, act=4, hidden=256, n_hidden=2, batch=256, steps=3, seed=4, full=False, head_scale=0.25),
    "humanoid_m1": dict(obs=348, act=17, hidden=256, n_hidden=2, batch=256, steps=2, seed=5, full=False, head_scale=0.25),
    "c2_humanoid_m2": dict(obs=348, act=17, hidden=512, n_hidden=3, batch=256, steps=2, seed=6, full=False, head_scale=0.25),
    "c3_nao_m2": dict(obs=661, act=23, hidden=512, n_hidden=3, batch=256, steps=1, seed=7, full=False, head_scale=0.25),
    "ckpt376_m1": dict(obs=376, act=17, hidden=256, n_hidden=2, batch=256, steps=1, seed=8, full=False, head_scale=0.25),
    # deliberately saturated (Xavier heads, |x_t| up to ~20): ill-conditioned in the reference itself, loose tolerance
    "humanoid_m1_saturated": dict(obs=348, act=17, hidden=256, n_hidden=2, batch=256, steps=1, seed=9, full=False, loose=True),
}

PER_CASES = {
    # n pushes into `capacity`; priorities then overwritten by the named distribution
    "small": dict(n=300, capacity=1000, batch=64, calls=3, seed=11, obs=3, act=2, dist="halfnormal"),
    "wrapped": dict(n=1500, capacity=1000, batch=256, calls=3, seed=12, obs=3, act=2, dist="halfnormal"),
    "floor": dict(n=20000, capacity=20000, batch=256, calls=4, seed=13, obs=2, act=1, dist="floor1pct"),
    "lognormal": dict(n=20000, capacity=20000, batch=256, calls=4, seed=14, obs=2, act=1, dist="lognormal3"),
    "fresh": dict(n=5000, capacity=8192, batch=256, calls=2, seed=15, obs=2, act=1, dist="fresh"),
    "short": dict(n=40, capacity=64, batch=256, calls=2, seed=16, obs=2, act=1, dist="halfnormal"),
}

UNIFORM_CASES = {
    "pool": dict(n=700, capacity=1000, batch=256, calls=3, seed=21),        # random.sample pool path (n <= setsize)
    "set": dict(n=5000, capacity=100000, batch=256, calls=3, seed=22),      # random.sample set/rejection path
    "wrapped": dict(n=2600, capacity=1000, batch=64, calls=3, seed=23),     # deque(maxlen) eviction
}


def per_transitions(case):
    rng = np.random.RandomState(500 + case["seed"])
    n = case["n"]
    return dict(
        s=rng.standard_normal((n, case["obs"])),                 # float64 like MuJoCo obs (walk_env.py:30)
        a=rng.uniform(-0.4, 0.4, (n, case["act"])).astype(np.float32),
        r=rng.standard_normal(n),
        s2=rng.standard_normal((n, case["obs"])),
        d=rng.uniform(size=n) < 0.05,
    )


def per_priorities(case):
    rng = np.random.RandomState(600 + case["seed"])
    m = min(case["n"], case["capacity"])
    dist = case["dist"]
    if dist == "halfnormal":
        p = np.abs(rng.standard_normal(m)) + 1e-6
    elif dist == "floor1pct":
        p = np.abs(rng.standard_normal(m)) + 1e-6
        p[rng.uniform(size=m) < 0.01] = 1e-6
    elif dist == "lognormal3":
        p = np.exp(3.0 * rng.standard_normal(m))
    elif dist == "fresh":
        p = np.ones(m)
    else:
        raise ValueError(dist)
    return p.astype(np.float32)


def per_td(case):
    rng = np.random.RandomState(700 + case["seed"])
    return np.abs(rng.standard_normal(case["batch"])).astype(np.float32)


# networks as callables (QNetwork.forward, GaussianPolicy.forward / .sample) on the initial weights of these update cases
NETS_CASES = ("tiny_m1", "tiny_m2", "c2_humanoid_m2")


def nets_inputs(case, n):
    rng = np.random.RandomState(900 + case["seed"])
    return dict(s=rng.standard_normal((n, case["obs"])).astype(np.float32),
                a=rng.uniform(-0.4, 0.4, (n, case["act"])).astype(np.float32),
                eps=rng.standard_normal((n, case["act"])).astype(np.float32))


def ckpt_transitions(case, n=20):
    """Transitions pushed into the buffer before a checkpoint is written (float64 observations like MuJoCo, python scalars)."""
    rng = np.random.RandomState(950 + case["seed"])
    return [(rng.standard_normal(case["obs"]), rng.uniform(-0.4, 0.4, case["act"]).astype(np.float32), float(rng.standard_normal()),
             rng.standard_normal(case["obs"]), bool(i % 6 == 5)) for i in range(n)]
