"""TEST INFRASTRUCTURE -- plain PyTorch (autograd + torch.optim.Adam) restatement of the reference's learner step.

Purpose: the like-for-like GPU baseline SURVEY 8d asks for ("reference-on-CUDA eager on the B200 box").  The reference itself
(/root/reference) cannot travel to the GPU box, so its update is restated here with stock torch ops, op for op in the order of
sac_imp.py:74-144, and PINNED to the same golden vectors as the numpy oracle (tests/test_oracle_torch_golden.py: losses of the
live reference on seeded inputs).  Run with device='cuda' it launches what the reference would launch: ~2100 ATen dispatches and
three `.item()` synchronisations per update.  Only tests/ and bench.py's baseline legs import this file; the product never does.

    networks     networks_model1.py:6-99 (2 hidden layers) / networks_model2.py:18-120 (3 hidden layers)
    update       sac_imp.py:74-144, Polyak :146-152
    PER (host)   replay_buffer.py:48-87 -- the reference keeps its buffers in host numpy whatever `device` is
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LOG_STD_MIN, LOG_STD_MAX = -20.0, 2.0


def _leaf(a, device):
    return torch.tensor(np.asarray(a, np.float32), device=device, requires_grad=True)


class TorchSAC:
    """Functional twin-critic SAC learner on dicts of leaf tensors keyed like the reference's state_dicts."""

    def __init__(self, state, device="cpu"):
        """state: oracle.sac_oracle_np.SACState (the weights / hyper-parameters every parity case starts from)."""
        self.device, self.n_hidden = torch.device(device), state.n_hidden
        self.gamma, self.tau, self.lr = state.gamma, state.tau, state.lr
        self.scale, self.bias = state.action_scale, state.action_bias
        self.auto = state.automatic_entropy_tuning
        self.nets = {n: {k: _leaf(v, self.device) for k, v in getattr(state, n).items()} for n in ("policy", "q1", "q2")}
        self.targets = {n: {k: torch.tensor(v, device=self.device) for k, v in getattr(state, n + "_target").items()} for n in ("q1", "q2")}
        self.opt = {n: torch.optim.Adam(list(p.values()), lr=self.lr) for n, p in self.nets.items()}      # sac_imp.py:39-41
        self.alpha = state.alpha                                                                           # python float (sac_imp.py:23)
        self.target_entropy = -float(state.act)                                                            # sac_imp.py:46
        self.log_alpha = torch.zeros(1, device=self.device, requires_grad=True)                            # sac_imp.py:48
        self.alpha_opt = torch.optim.Adam([self.log_alpha], lr=self.lr)

    # ---- networks -------------------------------------------------------------------------------------------------
    def _trunk(self, P, x):
        for i in range(1, self.n_hidden + 1):
            x = F.linear(x, P[f"fc{i}.weight"], P[f"fc{i}.bias"])
            if f"ln{i}.weight" in P:      # the product's opt-in LayerNorm variant (NOT in the reference): torch.nn.LayerNorm semantics, eps 1e-5
                x = F.layer_norm(x, (x.shape[-1],), P[f"ln{i}.weight"], P[f"ln{i}.bias"], 1e-5)
            x = F.relu(x)
        return x

    def q(self, P, s, a):                                    # networks_model1.py:27-33
        h = self._trunk(P, torch.cat([s, a], dim=1))
        k = self.n_hidden + 1
        return F.linear(h, P[f"fc{k}.weight"], P[f"fc{k}.bias"])

    def sample(self, s, eps=None):                           # networks_model1.py:65-99
        P = self.nets["policy"]
        h = self._trunk(P, s)
        mean = F.linear(h, P["mean.weight"], P["mean.bias"])
        log_std = torch.clamp(F.linear(h, P["log_std.weight"], P["log_std.bias"]), LOG_STD_MIN, LOG_STD_MAX)
        std = log_std.exp()
        if eps is None:
            eps = torch.randn_like(mean)
        x_t = mean + eps * std                               # Normal.rsample
        y_t = torch.tanh(x_t)
        action = y_t * self.scale + self.bias
        log_prob = -((x_t - mean) ** 2) / (2 * std * std) - log_std - math.log(math.sqrt(2 * math.pi))     # Normal.log_prob
        log_prob = log_prob - torch.log(self.scale * (1 - y_t.pow(2)) + 1e-6)
        return action, log_prob.sum(1, keepdim=True)

    # ---- one update (sac_imp.py:74-144) -------------------------------------------------------------------------------
    def update(self, s, a, r, s2, d, eps_next=None, eps_cur=None, weights=None):
        """s, a, r, s2, d: tensors on self.device (r, d as [B,1]); weights: optional IS weights [B,1] (extension H10)."""
        with torch.no_grad():
            a2, lp2 = self.sample(s2, eps_next)
            q_next = torch.min(self.q(self.targets["q1"], s2, a2), self.q(self.targets["q2"], s2, a2))
            y = r + (1 - d) * self.gamma * (q_next - self.alpha * lp2)
        q1, q2 = self.q(self.nets["q1"], s, a), self.q(self.nets["q2"], s, a)
        if weights is None:
            q1_loss, q2_loss = F.mse_loss(q1, y), F.mse_loss(q2, y)
        else:
            q1_loss, q2_loss = (weights * (q1 - y) ** 2).mean(), (weights * (q2 - y) ** 2).mean()
        td = (q1 - y).detach().abs()
        for n, loss in (("q1", q1_loss), ("q2", q2_loss)):
            self.opt[n].zero_grad()
            loss.backward()
            self.opt[n].step()
        a_new, lp = self.sample(s, eps_cur)
        q_new = torch.min(self.q(self.nets["q1"], s, a_new), self.q(self.nets["q2"], s, a_new))
        policy_loss = (self.alpha * lp - q_new).mean()
        self.opt["policy"].zero_grad()
        policy_loss.backward()
        self.opt["policy"].step()
        if self.auto:
            alpha_loss = -(self.log_alpha * (lp + self.target_entropy).detach()).mean()
            self.alpha_opt.zero_grad()
            alpha_loss.backward()
            self.alpha_opt.step()
            self.alpha = self.log_alpha.exp()
        with torch.no_grad():                                # sac_imp.py:146-152
            for n in ("q1", "q2"):
                for k, t in self.targets[n].items():
                    t.copy_(t * (1.0 - self.tau) + self.nets[n][k] * self.tau)
        return {"q1_loss": q1_loss.item(), "q2_loss": q2_loss.item(), "policy_loss": policy_loss.item()}, td

    def update_from_numpy(self, b, use_eps=True, weights=None):
        dev = self.device
        t = lambda x, col=False: torch.as_tensor(np.asarray(x, np.float32), device=dev).reshape(-1, 1) if col else torch.as_tensor(np.asarray(x, np.float32), device=dev)
        return self.update(t(b["s"]), t(b["a"]), t(b["r"], True), t(b["s2"]), t(b["d"], True),
                           t(b["eps_next"]) if use_eps else None, t(b["eps_cur"]) if use_eps else None,
                           None if weights is None else t(weights, True))


class NumpyPER:
    """The reference's prioritized buffer arithmetic on host numpy (replay_buffer.py:48-87): O(N) p**alpha, normalise,
    np.random.choice, IS weights; per-element priority write-back."""

    def __init__(self, priorities, alpha=0.6, beta_start=0.4, beta_frames=100000):
        self.priorities = np.asarray(priorities, np.float32).copy()
        self.alpha, self.beta_start, self.beta_frames, self.frame = alpha, beta_start, beta_frames, 1

    def sample(self, batch):
        beta = min(1.0, self.beta_start + self.frame * (1.0 - self.beta_start) / self.beta_frames)
        self.frame += 1
        probs = self.priorities ** self.alpha
        probs /= probs.sum()
        idx = np.random.choice(len(probs), batch, p=probs)
        w = (len(probs) * probs[idx]) ** (-beta)
        w /= w.max()
        return idx, w.astype(np.float32)

    def update_priorities(self, idx, td):
        for i, p in zip(idx, td):
            self.priorities[i] = p.item() + 1e-6
