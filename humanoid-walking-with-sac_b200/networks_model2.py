"""Three-hidden-layer ("enhanced") critic and actor: the reference's networks_model2.py surface
(QNetwork :18-46 Xavier init; GaussianPolicy :49-120 orthogonal init, extra `device` argument).
See networks_model1.py for how these modules alias the device arena."""
import torch
import torch.nn as nn

from . import networks_model1 as _m1

N_HIDDEN = 3
DEFAULT_HIDDEN = 512


def _init_orthogonal(module):
    # networks_model2.py:74-83
    if isinstance(module, nn.Linear):
        torch.nn.init.orthogonal_(module.weight, gain=1.0)
        torch.nn.init.constant_(module.bias, 0.0)


class QNetwork(_m1.QNetwork):
    N_HIDDEN = N_HIDDEN

    def __init__(self, state_dim, action_dim, hidden_dim=DEFAULT_HIDDEN, *, layer_norm=False):
        super().__init__(state_dim, action_dim, hidden_dim, layer_norm=layer_norm)


class GaussianPolicy(_m1.GaussianPolicy):
    N_HIDDEN = N_HIDDEN
    _init_weights = staticmethod(_init_orthogonal)

    def __init__(self, state_dim, action_dim, hidden_dim=DEFAULT_HIDDEN, device="cuda", action_bounds=None, *, layer_norm=False):
        super().__init__(state_dim, action_dim, hidden_dim, action_bounds, layer_norm=layer_norm)
        self.device = device     # kept for signature compatibility; placement is decided by the owning SAC
