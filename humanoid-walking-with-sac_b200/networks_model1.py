"""Two-hidden-layer critic and tanh-Gaussian actor: same names, constructor arguments, parameter names,
shapes and initialisers as the reference's networks_model1.py (QNetwork :6-33, GaussianPolicy :36-99).

Here the modules are the *state_dict face* of the learner: once a `SAC` owns them, every parameter
aliases a slice of the device arena that the fused update kernels read and write, so `state_dict()`,
`load_state_dict()` and `torch.save` keep the reference layout while no autograd graph is ever
built.  `forward` / `sample` evaluate rows through the library (sacb_q_forward / sacb_policy_forward).
"""
import numpy as np
import torch
import torch.nn as nn

from . import _native as N

N_HIDDEN = 2
DEFAULT_HIDDEN = 256
LOG_STD_MIN, LOG_STD_MAX = -20.0, 2.0


def _init_xavier(module):
    # networks_model1.py:22-25 / :60-63 -- same torch calls so a seeded construction consumes the global
    # generator exactly as the reference does (bit-identical initial weights for a given torch.manual_seed)
    if isinstance(module, nn.Linear):
        torch.nn.init.xavier_uniform_(module.weight)
        torch.nn.init.constant_(module.bias, 0)


class _ArenaModule(nn.Module):
    """nn.Module whose parameters can be re-pointed at device memory owned by a SAC handle."""

    N_HIDDEN = N_HIDDEN

    def _bind(self, owner, net_id):
        """Upload the current (host) parameter values into the arena and alias them from now on."""
        lib = N.lib()
        self._owner, self._net_id = owner, net_id
        for t, (name, p) in enumerate(list(self.named_parameters())):
            host = N.f32(p.detach().cpu().numpy())
            N.check(lib.sacb_import_tensor(owner._h, 0, net_id, N.SLOT_PARAM, t, N.ptr(host), host.size))
            alias = owner._param_alias(net_id, t, p.shape)
            mod_name, _, leaf = name.rpartition(".")
            getattr(self, mod_name)._parameters[leaf] = nn.Parameter(alias, requires_grad=False)
        return self

    def _bound(self):
        return getattr(self, "_owner", None) is not None


class QNetwork(_ArenaModule):
    def __init__(self, state_dim, action_dim, hidden_dim=DEFAULT_HIDDEN, *, layer_norm=False):
        """layer_norm=True (extension, default off: the reference has no LayerNorm -- networks_model2.py:86 is only a comment) puts
        `ln{i}` = nn.LayerNorm(hidden_dim) between `fc{i}` and its ReLU; parameters then come as fc1.*, ln1.*, fc2.*, ln2.*, ..."""
        super().__init__()
        widths = [state_dim + action_dim] + [hidden_dim] * self.N_HIDDEN
        for i in range(self.N_HIDDEN):
            setattr(self, f"fc{i + 1}", nn.Linear(widths[i], widths[i + 1]))
            if layer_norm:
                setattr(self, f"ln{i + 1}", nn.LayerNorm(hidden_dim))
        setattr(self, f"fc{self.N_HIDDEN + 1}", nn.Linear(hidden_dim, 1))
        self.apply(self._init_weights)

    _init_weights = staticmethod(_init_xavier)

    def forward(self, state, action):
        if not self._bound():
            raise RuntimeError("QNetwork is evaluated by the CUDA library: construct it through SAC (no PyTorch fallback)")
        s, a = N.f32(torch.as_tensor(state).detach().cpu().numpy()), N.f32(torch.as_tensor(action).detach().cpu().numpy())
        s, a = s.reshape(-1, s.shape[-1]), a.reshape(-1, a.shape[-1])
        q = np.empty(s.shape[0], np.float32)
        N.check(N.lib().sacb_q_forward(self._owner._h, 0, self._net_id, N.ptr(s), N.ptr(a), s.shape[0], N.ptr(q)))
        return torch.from_numpy(q).unsqueeze(-1)


class GaussianPolicy(_ArenaModule):
    def __init__(self, state_dim, action_dim, hidden_dim=DEFAULT_HIDDEN, action_bounds=None, *, layer_norm=False):
        super().__init__()
        widths = [state_dim] + [hidden_dim] * self.N_HIDDEN
        for i in range(self.N_HIDDEN):
            setattr(self, f"fc{i + 1}", nn.Linear(widths[i], widths[i + 1]))
            if layer_norm:
                setattr(self, f"ln{i + 1}", nn.LayerNorm(hidden_dim))
        self.mean = nn.Linear(hidden_dim, action_dim)
        self.log_std = nn.Linear(hidden_dim, action_dim)
        lo, hi = (-0.4, 0.4) if action_bounds is None else action_bounds      # networks_model1.py:52-55
        self.action_scale = (hi - lo) / 2
        self.action_bias = (hi + lo) / 2
        self.apply(self._init_weights)

    _init_weights = staticmethod(_init_xavier)

    def forward(self, state):
        if not self._bound():
            raise RuntimeError("GaussianPolicy is evaluated by the CUDA library: construct it through SAC (no PyTorch fallback)")
        s = N.f32(torch.as_tensor(state).detach().cpu().numpy())
        s = s.reshape(-1, s.shape[-1])
        act = self.mean.out_features
        mean, log_std = np.empty((s.shape[0], act), np.float32), np.empty((s.shape[0], act), np.float32)
        N.check(N.lib().sacb_policy_forward(self._owner._h, 0, N.ptr(s), s.shape[0], N.ptr(mean), N.ptr(log_std)))
        return torch.from_numpy(mean), torch.from_numpy(log_std)

    def sample(self, state, eps=None):
        """networks_model1.py:78-99 evaluated by the library (sacb_policy_sample: the same device arithmetic as the update's
        sampling stage).  eps [n, act] replaces the N(0,1) draw of Normal.rsample; None draws on the device (Philox)."""
        if not self._bound():
            raise RuntimeError("GaussianPolicy is evaluated by the CUDA library: construct it through SAC (no PyTorch fallback)")
        s = N.f32(torch.as_tensor(state).detach().cpu().numpy())
        s = s.reshape(-1, s.shape[-1])
        act = self.mean.out_features
        e = None if eps is None else N.f32(torch.as_tensor(eps).detach().cpu().numpy()).reshape(s.shape[0], act)
        action, logp = np.empty((s.shape[0], act), np.float32), np.empty(s.shape[0], np.float32)
        N.check(N.lib().sacb_policy_sample(self._owner._h, 0, N.ptr(s), s.shape[0], N.ptr(e), N.ptr(action), N.ptr(logp)))
        return torch.from_numpy(action), torch.from_numpy(logp).unsqueeze(-1)
