#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the LIVE reference.

Runs only in the build container (needs /root/reference, torch CPU).  The GPU box never runs this;
it only reads the committed .npz files.  Inputs (weights, minibatches, eps draws, priorities) are
NOT stored: they regenerate from numpy's frozen legacy RandomState streams through
oracle/sac_oracle_np.py::make_state / make_batch, so the fixtures hold outputs only (plus the
PER p**alpha tables, because numpy's float32 pow is CPU-ISA dependent).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz
"""
import functools
import os
import random
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import numpy as np
import torch

import networks_model1
import networks_model2
import replay_buffer as ref_rb
import sac_imp

from oracle import sac_oracle_np as O
from tests.golden import cases

torch.set_num_threads(4)


def build_reference_agent(case, st):
    """Reference SAC on CPU with the numpy-initialised weights of `st` loaded (sac_imp.py:9-52)."""
    if case["n_hidden"] == 2:
        sac_imp.QNetwork = networks_model1.QNetwork
        sac_imp.GaussianPolicy = networks_model1.GaussianPolicy
    else:
        sac_imp.QNetwork = networks_model2.QNetwork
        sac_imp.GaussianPolicy = functools.partial(networks_model2.GaussianPolicy, device="cpu")
    agent = sac_imp.SAC(case["obs"], case["act"], hidden_dim=case["hidden"], device="cpu",
                        automatic_entropy_tuning=case.get("auto_entropy", True))
    for net, P in (("policy", st.policy), ("q1", st.q1), ("q2", st.q2),
                   ("q1_target", st.q1_target), ("q2_target", st.q2_target)):
        getattr(agent, net).load_state_dict({k: torch.from_numpy(v.copy()) for k, v in P.items()})
    return agent


class EpsInjector:
    """Replaces the N(0,1) draw inside Normal.rsample (torch/distributions/normal.py) by a queue."""

    def __init__(self):
        self.queue = []
        self._orig = torch.distributions.normal._standard_normal

    def __enter__(self):
        def fake(shape, dtype, device):
            e = self.queue.pop(0)
            assert tuple(e.shape) == tuple(shape), (e.shape, shape)
            return torch.from_numpy(np.ascontiguousarray(e)).to(dtype)
        torch.distributions.normal._standard_normal = fake
        return self

    def __exit__(self, *a):
        torch.distributions.normal._standard_normal = self._orig


def snapshot_grads_before_step(opt, store, key):
    orig = opt.step

    def step(*a, **k):
        store[key] = [p.grad.detach().clone().numpy() for g in opt.param_groups for p in g["params"]]
        return orig(*a, **k)
    opt.step = step


def summarize(arr):
    a = np.asarray(arr, np.float64).ravel()
    return np.array([a.sum(), np.sqrt((a * a).sum()), np.abs(a).max()])


def run_update_case(case):
    st = O.make_state(case["obs"], case["act"], case["hidden"], case["n_hidden"], seed=case["seed"],
                      bias_scale=case.get("bias_scale", 0.0), head_scale=case.get("head_scale", 1.0),
                      automatic_entropy_tuning=case.get("auto_entropy", True))
    agent = build_reference_agent(case, st)
    out = {}
    grads = {}
    snapshot_grads_before_step(agent.q1_optimizer, grads, "q1")
    snapshot_grads_before_step(agent.q2_optimizer, grads, "q2")
    snapshot_grads_before_step(agent.policy_optimizer, grads, "policy")
    if case.get("auto_entropy", True):
        snapshot_grads_before_step(agent.alpha_optimizer, grads, "log_alpha")
    losses = []
    alphas = []
    with EpsInjector() as inj:
        for step in range(case["steps"]):
            b = O.make_batch(case["obs"], case["act"], case["batch"], seed=case["seed"] * 100 + step)
            agent.replay_buffer.sample = lambda n, b=b: (b["s"], b["a"], b["r"], b["s2"], b["d"])
            inj.queue += [b["eps_next"], b["eps_cur"]]
            info = agent.update_parameters(case["batch"])
            losses.append([info["q1_loss"], info["q2_loss"], info["policy_loss"]])
            alphas.append(float(agent.alpha))
            if step == 0:
                qn = O.q_param_names(case["n_hidden"])
                pn = O.policy_param_names(case["n_hidden"])
                for net, names in (("q1", qn), ("q2", qn), ("policy", pn)):
                    for nm, g in zip(names, grads[net]):
                        if case["full"]:
                            out[f"grad/{net}/{nm}"] = g
                        out[f"gradsum/{net}/{nm}"] = summarize(g)
                if "log_alpha" in grads:
                    out["grad/log_alpha"] = grads["log_alpha"][0]
    out["losses"] = np.array(losses, np.float64)
    out["alphas"] = np.array(alphas, np.float64)
    if case.get("auto_entropy", True):
        out["log_alpha"] = agent.log_alpha.detach().numpy().copy()
    opts = {"policy": agent.policy_optimizer, "q1": agent.q1_optimizer, "q2": agent.q2_optimizer}
    for net in ("policy", "q1", "q2", "q1_target", "q2_target"):
        sd = getattr(agent, net).state_dict()
        for i, (nm, t) in enumerate(sd.items()):
            t = t.numpy()
            if case["full"]:
                out[f"param/{net}/{nm}"] = t.copy()
            out[f"paramsum/{net}/{nm}"] = summarize(t)
            out[f"paramhead/{net}/{nm}"] = t.ravel()[:64].copy()
            if net in opts and case["full"]:
                stt = opts[net].state_dict()["state"][i]
                out[f"adam_m/{net}/{nm}"] = stt["exp_avg"].numpy().copy()
                out[f"adam_v/{net}/{nm}"] = stt["exp_avg_sq"].numpy().copy()
    # select_action on the post-update policy (sac_imp.py:54-72)
    obs_vec = np.random.RandomState(77 + case["seed"]).standard_normal(case["obs"]).astype(np.float32)
    eps_vec = np.random.RandomState(78 + case["seed"]).standard_normal((1, case["act"])).astype(np.float32)
    out["select/eval"] = agent.select_action(obs_vec, evaluate=True)
    with EpsInjector() as inj:
        inj.queue.append(eps_vec)
        out["select/sample"] = agent.select_action(obs_vec, evaluate=False)
    return out


def run_per_case(case):
    """Live PrioritizedReplayBuffer (replay_buffer.py:25-90): push n, overwrite priorities, sample, update."""
    n, cap, B = case["n"], case["capacity"], case["batch"]
    buf = ref_rb.PrioritizedReplayBuffer(cap)
    tr = cases.per_transitions(case)
    for i in range(n):
        buf.push(tr["s"][i], tr["a"][i], tr["r"][i], tr["s2"][i], bool(tr["d"][i]))
    out = {"prio_after_push": buf.priorities.copy(), "pos_after_push": np.array(buf.pos)}
    pri = cases.per_priorities(case)
    m = min(n, cap)
    buf.priorities[:m] = pri[:m]
    out["p_alpha"] = (buf.priorities[:m] ** buf.alpha).astype(np.float32)     # replay_buffer.py:60 on THIS box
    idx_all, w_all, beta_all = [], [], []
    for call in range(case["calls"]):
        np.random.seed(case["seed"] * 10 + call)
        frame = buf.frame
        s, a, r, s2, d, idx, w = buf.sample(B)
        idx_all.append(idx.astype(np.int64))
        w_all.append(w.astype(np.float32))
        beta_all.append(min(1.0, buf.beta_start + frame * (1.0 - buf.beta_start) / buf.beta_frames))
        if call == 0:
            out["sample0/r"] = r
            out["sample0/s"] = s
            out["sample0/d"] = d
    out["idx"] = np.stack(idx_all)
    out["weights"] = np.stack(w_all)
    out["beta"] = np.array(beta_all)
    # update_priorities (replay_buffer.py:84-87) with duplicates inside the batch: last one wins
    td = cases.per_td(case)
    upd_idx = out["idx"][0].copy()
    upd_idx[1::7] = upd_idx[0]
    buf.update_priorities(upd_idx, torch.from_numpy(td))
    out["upd_idx"] = upd_idx
    out["prio_after_update"] = buf.priorities.copy()
    buf.push(tr["s"][0], tr["a"][0], tr["r"][0], tr["s2"][0], False)   # push after update: max-priority rule :38
    out["prio_after_push2"] = buf.priorities.copy()
    out["pos_after_push2"] = np.array(buf.pos)
    return out


def run_uniform_case(case):
    """Live ReplayBuffer (replay_buffer.py:5-22): deque(maxlen) ring + random.sample without replacement."""
    buf = ref_rb.ReplayBuffer(case["capacity"])
    for i in range(case["n"]):
        buf.push(np.full(3, i, np.float32), np.full(2, -i, np.float32), float(i), np.full(3, i + 0.5, np.float32), i % 5 == 0)
    out = {}
    random.seed(case["seed"])
    picks = []
    for call in range(case["calls"]):
        s, a, r, s2, d = buf.sample(case["batch"])
        picks.append(r.astype(np.int64))
        if call == 0:
            out["s"], out["a"], out["s2"], out["d"] = s, a, s2, d
    out["r_ids"] = np.stack(picks)
    out["len"] = np.array(len(buf))
    return out


def run_nets_case(case, n=32):
    """The networks as callables on the INITIAL weights: QNetwork.forward (networks_model1.py:27-33 / networks_model2.py:37-46),
    GaussianPolicy.forward (:65-76 / :85-97) and GaussianPolicy.sample (:78-99 / :99-120) with the rsample draw injected."""
    st = O.make_state(case["obs"], case["act"], case["hidden"], case["n_hidden"], seed=case["seed"],
                      bias_scale=case.get("bias_scale", 0.0), head_scale=case.get("head_scale", 1.0))
    agent = build_reference_agent(case, st)
    inp = cases.nets_inputs(case, n)
    s, a = torch.from_numpy(inp["s"]), torch.from_numpy(inp["a"])
    out = {}
    with torch.no_grad():
        out["q1"] = agent.q1(s, a).numpy()
        out["q2_target"] = agent.q2_target(s, a).numpy()
        mean, log_std = agent.policy(s)
        out["mean"], out["log_std"] = mean.numpy(), log_std.numpy()
        with EpsInjector() as inj:
            inj.queue.append(inp["eps"])
            action, logp = agent.policy.sample(s)
        out["action"], out["log_prob"] = action.numpy(), logp.numpy()
    return out


def _seeded_updates(agent, case, steps, first=0):
    losses = []
    with EpsInjector() as inj:
        for step in range(first, first + steps):
            b = O.make_batch(case["obs"], case["act"], case["batch"], seed=case["seed"] * 100 + step)
            agent.replay_buffer.sample = lambda n, b=b: (b["s"], b["a"], b["r"], b["s2"], b["d"])
            inj.queue += [b["eps_next"], b["eps_cur"]]
            info = agent.update_parameters(case["batch"])
            losses.append([info["q1_loss"], info["q2_loss"], info["policy_loss"]])
    return np.array(losses, np.float64)


def run_ckpt_written_by_reference():
    """Files written by the reference's own save() (sac_imp.py:154-162) and save_checkpoint() (:177-201) after two seeded updates
    of tiny_m1 with 20 transitions in the deque, plus what the reference itself does NEXT (third update, select_action)."""
    case = cases.UPDATE_CASES["tiny_m1"]
    st = O.make_state(case["obs"], case["act"], case["hidden"], case["n_hidden"], seed=case["seed"], bias_scale=case.get("bias_scale", 0.0))
    agent = build_reference_agent(case, st)
    sample_fn = agent.replay_buffer.sample
    _seeded_updates(agent, case, 2)
    for t in cases.ckpt_transitions(case):
        agent.replay_buffer.push(*t)
    agent.replay_buffer.sample = sample_fn
    agent.save(os.path.join(HERE, "ckpt_ref_save.pt"))
    agent.save_checkpoint(os.path.join(HERE, "ckpt_ref_checkpoint.pt"), episode=7, total_steps=123)
    out = {"next_losses": _seeded_updates(agent, case, 1, first=2), "next_alpha": np.array(float(agent.alpha))}
    obs_vec = np.random.RandomState(77 + case["seed"]).standard_normal(case["obs"]).astype(np.float32)
    out["select_eval_after"] = agent.select_action(obs_vec, evaluate=True)
    return out


SHIPPED = {"bipedal": ("sac_BipedalWalker-v3_1737453113", dict(obs=24, act=4, hidden=256, n_hidden=2, batch=256, seed=31)),
           "humanoid376": ("sac_Humanoid-v5_1734629000", dict(obs=376, act=17, hidden=256, n_hidden=2, batch=256, seed=32))}


def run_shipped_checkpoint(tag):
    """The reference loads its own shipped results/*/best_model.pt (sac_imp.py:164-173), acts, and takes one seeded update."""
    import shutil
    dirname, case = SHIPPED[tag]
    src = os.path.join("/root/reference/results", dirname, "best_model.pt")
    shutil.copyfile(src, os.path.join(HERE, f"shipped_{tag}_best_model.pt"))      # reference-held fixture (binary artefact, not source)
    sac_imp.QNetwork, sac_imp.GaussianPolicy = networks_model1.QNetwork, networks_model1.GaussianPolicy
    agent = sac_imp.SAC(case["obs"], case["act"], hidden_dim=case["hidden"], device="cpu")
    # sac_imp.py:164-173 verbatim, except map_location: the shipped files hold CUDA tensors and this container has no GPU
    ck = torch.load(src, map_location="cpu", weights_only=False)
    for net in ("policy", "q1", "q2", "q1_target", "q2_target"):
        getattr(agent, net).load_state_dict(ck[f"{net}_state_dict"])
    agent.alpha = ck["alpha"]
    out = {"alpha_loaded": np.array(float(agent.alpha))}
    obs_mat = np.random.RandomState(77 + case["seed"]).standard_normal((8, case["obs"])).astype(np.float32)
    out["select_eval"] = np.stack([agent.select_action(o, evaluate=True) for o in obs_mat])
    inp = cases.nets_inputs(case, 16)
    with torch.no_grad():
        out["q1"] = agent.q1(torch.from_numpy(inp["s"]), torch.from_numpy(inp["a"])).numpy()
    out["losses"] = _seeded_updates(agent, case, 1)
    out["alpha_after"] = np.array(float(agent.alpha))
    out["select_eval_after"] = np.stack([agent.select_action(o, evaluate=True) for o in obs_mat])
    return out


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else None
    if only in (None, "nets"):
        for name in cases.NETS_CASES:
            out = run_nets_case(cases.UPDATE_CASES[name])
            np.savez_compressed(os.path.join(HERE, f"nets_{name}.npz"), **out)
            print("nets", name, out["q1"][:2].ravel(), out["log_prob"][:2].ravel())
    if only in (None, "ckpt"):
        out = run_ckpt_written_by_reference()
        np.savez_compressed(os.path.join(HERE, "ckpt_ref_expected.npz"), **out)
        print("ckpt written by the reference; next losses", out["next_losses"])
        for tag in SHIPPED:
            out = run_shipped_checkpoint(tag)
            np.savez_compressed(os.path.join(HERE, f"shipped_{tag}_expected.npz"), **out)
            print("shipped", tag, "alpha", out["alpha_loaded"], "losses", out["losses"])
    if only is not None:
        return
    for name, case in cases.UPDATE_CASES.items():
        out = run_update_case(case)
        np.savez_compressed(os.path.join(HERE, f"update_{name}.npz"), **out)
        print("update", name, {k: v for k, v in zip(("q1", "q2", "pi"), out["losses"][-1])}, "alpha", out["alphas"][-1])
    for name, case in cases.PER_CASES.items():
        out = run_per_case(case)
        np.savez_compressed(os.path.join(HERE, f"per_{name}.npz"), **out)
        print("per", name, out["idx"][0][:6], out["weights"][0][:3])
    for name, case in cases.UNIFORM_CASES.items():
        out = run_uniform_case(case)
        np.savez_compressed(os.path.join(HERE, f"uniform_{name}.npz"), **out)
        print("uniform", name, out["r_ids"][0][:6])


if __name__ == "__main__":
    main()
