"""B200-native SAC learner hot path: drop-in for the reference's sac_imp / replay_buffer / networks_model* modules.

    import humanoid_walking_with_sac_b200 as hw
    hw.install()                     # `from sac_imp import SAC` (trainer.py:5) now resolves to this package
    agent = hw.SAC(348, 17, 256)

The arithmetic runs in libsacb200.so (hand-written sm_100a CUDA, C ABI in include/sacb200.h); importing
the package does not need a GPU, constructing a SAC / buffer does.  There is no CPU fallback.
"""
import sys

from . import _native, distributed, networks_model1, networks_model2, replay_buffer, sac_imp
from .replay_buffer import PrioritizedReplayBuffer, ReplayBuffer
from .sac_imp import SAC, PopulationSAC

__all__ = ["SAC", "PopulationSAC", "ReplayBuffer", "PrioritizedReplayBuffer", "install", "use_networks", "networks_model1", "networks_model2", "distributed"]


def use_networks(variant):
    """Equivalent of editing sac_imp.py:4: variant 'model1' (2x hidden) or 'model2' (3x hidden, orthogonal policy init)."""
    mod = {"model1": networks_model1, "model2": networks_model2}[variant]
    sac_imp.QNetwork, sac_imp.GaussianPolicy = mod.QNetwork, mod.GaussianPolicy


def install():
    """Register this package's modules under the reference's top-level module names."""
    for name, mod in (("sac_imp", sac_imp), ("replay_buffer", replay_buffer), ("networks_model1", networks_model1),
                      ("networks_model2", networks_model2)):
        sys.modules[name] = mod
