"""Soft Actor-Critic learner behind the reference's sac_imp.SAC API, executed by hand-written sm_100a kernels.

Drop-in surface (reference sac_imp.py): constructor :9-52, select_action :54-72, update_parameters :74-144,
save/load :154-173, save_checkpoint/load_checkpoint :177-233, attributes policy / q1 / q2 / q1_target /
q2_target / *_optimizer / alpha / log_alpha / replay_buffer / gamma / tau / device / target_entropy.

Which networks are used follows the module-level names `QNetwork` / `GaussianPolicy`, exactly like the
reference's `from networks_model1 import ...` (sac_imp.py:4): rebind them to networks_model2's classes to
get the 3x512 variant.  Keyword-only arguments after `device` are extensions (replay kind, math mode...).
"""
import ctypes
import random

import numpy as np
import torch

from . import _native as N
from .networks_model1 import GaussianPolicy, QNetwork
from .replay_buffer import PrioritizedReplayBuffer, ReplayBuffer

_NETS = ("policy", "q1", "q2", "q1_target", "q2_target")


class ArenaAdam:
    """torch.optim.Adam look-alike over the device arena: the step itself is fused into the update kernels;
    this object only serves `state_dict()` / `load_state_dict()` in the reference checkpoint layout."""

    def __init__(self, owner, net_id, module, lr):
        self._owner, self._net_id, self._module = owner, net_id, module
        n = len(list(module.parameters())) if module is not None else 1
        self.param_groups = [dict(lr=lr, betas=(0.9, 0.999), eps=1e-08, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                                  capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False, params=list(range(n)))]

    def _step_count(self):
        sc = self._owner._scalars()
        return {N.NET_POLICY: sc.step_policy, N.NET_Q1: sc.step_q1, N.NET_Q2: sc.step_q2, None: sc.step_alpha}[self._net_id]

    def _set_step(self, v):
        sc = self._owner._scalars()
        name = {N.NET_POLICY: "step_policy", N.NET_Q1: "step_q1", N.NET_Q2: "step_q2", None: "step_alpha"}[self._net_id]
        setattr(sc, name, int(v))
        self._owner._set_scalars(sc)

    def zero_grad(self, set_to_none=True):
        pass

    def step(self):
        raise RuntimeError("the optimizer step is fused into SAC.update_parameters")

    def state_dict(self):
        step = self._step_count()
        state = {}
        if step > 0:
            if self._net_id is None:
                sc = self._owner._scalars()
                state[0] = {"step": torch.tensor(float(step)), "exp_avg": torch.tensor([sc.log_alpha_m]), "exp_avg_sq": torch.tensor([sc.log_alpha_v])}
            else:
                for i, p in enumerate(self._module.parameters()):
                    state[i] = {"step": torch.tensor(float(step)),
                                "exp_avg": self._owner._tensor(self._net_id, N.SLOT_ADAM_M, i, p.shape),
                                "exp_avg_sq": self._owner._tensor(self._net_id, N.SLOT_ADAM_V, i, p.shape)}
        return {"state": state, "param_groups": [dict(g) for g in self.param_groups]}

    def load_state_dict(self, sd):
        state = sd["state"]
        if not state:
            self._set_step(0)
            self._restore_lr(sd)
            return
        first = state[min(state)]
        self._set_step(int(float(first["step"])))
        if self._net_id is None:
            sc = self._owner._scalars()
            sc.log_alpha_m = float(torch.as_tensor(first["exp_avg"]).reshape(-1)[0])
            sc.log_alpha_v = float(torch.as_tensor(first["exp_avg_sq"]).reshape(-1)[0])
            self._owner._set_scalars(sc)
            self._restore_lr(sd)
            return
        for i, st in state.items():
            for slot, key in ((N.SLOT_ADAM_M, "exp_avg"), (N.SLOT_ADAM_V, "exp_avg_sq")):
                host = N.f32(torch.as_tensor(st[key]).detach().cpu().numpy())
                N.check(N.lib().sacb_import_tensor(self._owner._h, 0, self._net_id, slot, int(i), N.ptr(host), host.size))
        self._restore_lr(sd)

    def _restore_lr(self, sd):
        # torch.optim.Adam.load_state_dict restores the learning rate of the checkpoint (sac_imp.py:215-218); the device
        # Adam of all four optimizers shares ONE rate (sac_imp.py:39-49), so a differing value re-derives its step-size table
        if sd.get("param_groups") and "lr" in sd["param_groups"][0]:
            lr = float(sd["param_groups"][0]["lr"])
            self._owner._set_lr(lr)


class SAC:
    """Soft Actor-Critic for continuous actions; one fused device program per update_parameters call."""

    def __init__(self, state_dim, action_dim, hidden_dim=256, gamma=0.99, tau=0.005, lr=3e-4, alpha=0.2,
                 automatic_entropy_tuning=True, device="cuda" if torch.cuda.is_available() else "cpu", *,
                 replay="uniform", capacity=1000000, max_batch=256, math="bf16x3", launch="staged", seed=None,
                 per_alpha=0.6, per_beta_start=0.4, per_beta_frames=100000, per_weighted_loss=False, action_bounds=None, layer_norm=False):
        if not str(device).startswith("cuda"):
            raise RuntimeError("this SAC runs on a B200 only (device='cuda[:i]'); there is no CPU path")
        self.gamma, self.tau, self.device = gamma, tau, device
        self.automatic_entropy_tuning = automatic_entropy_tuning
        self._lr = lr
        dev_index = torch.device(device).index or 0

        # networks: built on the host with the reference's initialisers (same global-RNG consumption order as
        # sac_imp.py:28-36), then uploaded; afterwards every parameter aliases the device arena
        kw = {} if action_bounds is None else {"action_bounds": action_bounds}
        ln = {"layer_norm": True} if layer_norm else {}      # extension, default off (no LayerNorm in the reference): DESIGN.md
        self.layer_norm = bool(layer_norm)
        self.policy = GaussianPolicy(state_dim, action_dim, hidden_dim, **kw, **ln)
        self.q1 = QNetwork(state_dim, action_dim, hidden_dim, **ln)
        self.q2 = QNetwork(state_dim, action_dim, hidden_dim, **ln)
        self.q1_target = QNetwork(state_dim, action_dim, hidden_dim, **ln)
        self.q2_target = QNetwork(state_dim, action_dim, hidden_dim, **ln)
        self.q1_target.load_state_dict(self.q1.state_dict())
        self.q2_target.load_state_dict(self.q2.state_dict())
        n_hidden = getattr(self.q1, "N_HIDDEN", 2)
        if getattr(self.policy, "N_HIDDEN", n_hidden) != n_hidden:
            raise ValueError("QNetwork and GaussianPolicy must come from the same networks_model* variant")

        cfg = N.default_config()
        cfg.obs_dim, cfg.act_dim, cfg.hidden_dim, cfg.n_hidden = state_dim, action_dim, hidden_dim, n_hidden
        cfg.gamma, cfg.tau, cfg.lr, cfg.alpha0, cfg.auto_entropy = gamma, tau, lr, alpha, int(bool(automatic_entropy_tuning))
        cfg.action_scale, cfg.action_bias = self.policy.action_scale, self.policy.action_bias
        cfg.replay_kind = N.REPLAY_PER if replay == "per" else N.REPLAY_UNIFORM
        cfg.capacity, cfg.max_batch, cfg.n_agents, cfg.device = capacity, max_batch, 1, dev_index
        cfg.per_alpha, cfg.per_beta_start, cfg.per_beta_frames = per_alpha, per_beta_start, per_beta_frames
        cfg.per_weighted_loss = int(bool(per_weighted_loss))
        cfg.layer_norm = int(bool(layer_norm))
        cfg.math_mode = {"fp32": N.MATH_FP32, "bf16x3": N.MATH_BF16X3}[math]
        cfg.launch_mode = {"staged": N.LAUNCH_STAGED, "persistent": N.LAUNCH_PERSISTENT}[launch]
        cfg.seed = random.getrandbits(63) if seed is None else int(seed)
        self._cfg = cfg
        self._h = N.create(cfg)
        self._arena_base = None
        for net_id, name in enumerate(_NETS):
            getattr(self, name)._bind(self, net_id)

        self.policy_optimizer = ArenaAdam(self, N.NET_POLICY, self.policy, lr)
        self.q1_optimizer = ArenaAdam(self, N.NET_Q1, self.q1, lr)
        self.q2_optimizer = ArenaAdam(self, N.NET_Q2, self.q2, lr)
        if automatic_entropy_tuning:
            self.target_entropy = -action_dim                                    # sac_imp.py:46
            self.alpha_optimizer = ArenaAdam(self, None, None, lr)

        self.replay_buffer = (PrioritizedReplayBuffer(capacity, per_alpha, per_beta_start, per_beta_frames)
                              if replay == "per" else ReplayBuffer(capacity))
        self.replay_buffer._bind(self)
        self._alpha_is_float = True        # python float until the first update (quirk Q1)
        self._alias_version = None

    def _param_alias(self, net_id, t, shape):
        """Parameter tensor `t` of a network as a VIEW of one tensor spanning the parameter arena: every alias then shares that
        tensor's version counter, and one integer tells whether anything was written through any of them."""
        lib = N.lib()
        if self._arena_base is None:
            lo, hi = None, 0
            for net in range(len(_NETS)):
                for k in range(lib.sacb_num_tensors(self._h, net)):
                    rows, cols, off, dev = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_void_p()
                    N.check(lib.sacb_tensor_info(self._h, net, k, ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(off)))
                    N.check(lib.sacb_tensor_dev(self._h, 0, net, N.SLOT_PARAM, k, ctypes.byref(dev)))
                    lo = dev.value if lo is None else min(lo, dev.value)
                    hi = max(hi, dev.value + 4 * rows.value * cols.value)
            self._arena_base = (lo, torch.as_tensor(N.DevArray(lo, ((hi - lo) // 4,), self), device=f"cuda:{self._cfg.device}"))
        dev = ctypes.c_void_p()
        N.check(lib.sacb_tensor_dev(self._h, 0, net_id, N.SLOT_PARAM, t, ctypes.byref(dev)))
        lo, base = self._arena_base
        first, n = (dev.value - lo) // 4, int(np.prod(shape))
        return base[first:first + n].view(tuple(shape))

    def _publish_alias_writes(self):
        """The GEMMs read bf16 hi/lo shadows that the update's own epilogues keep current.  A write through the torch aliases
        (load_state_dict, `with torch.no_grad(): p.mul_(...)`, ...) bumps the version counter the aliases share (they are views of
        one tensor): the library is then told to re-derive every shadow before the next update.  (`p.data` views have their own
        counter: call invalidate_shadows() after writing through one.)"""
        v = self._arena_base[1]._version
        if v != self._alias_version:
            self._alias_version = v
            N.check(N.lib().sacb_invalidate_shadows(self._h))

    def invalidate_shadows(self):
        N.check(N.lib().sacb_invalidate_shadows(self._h))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None:
            try:
                N.lib().sacb_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- scalars ---------------------------------------------------------------------------------------------
    def _scalars(self):
        sc = N.Scalars()
        N.check(N.lib().sacb_get_scalars(self._h, 0, ctypes.byref(sc)))
        return sc

    def _set_scalars(self, sc):
        N.check(N.lib().sacb_set_scalars(self._h, 0, ctypes.byref(sc)))

    def _set_lr(self, lr):
        N.check(N.lib().sacb_set_lr(self._h, lr))
        self._lr = lr
        for opt in ("policy_optimizer", "q1_optimizer", "q2_optimizer", "alpha_optimizer"):
            if hasattr(self, opt):
                getattr(self, opt).param_groups[0]["lr"] = lr

    def _library_stream(self):
        """The handle's private CUDA stream as a torch stream: torch-side reads / writes of the aliased arena (state_dict,
        load_state_dict, torch.save) are enqueued on it, i.e. ordered against the kernels of the library."""
        if getattr(self, "_ext_stream", None) is None:
            sp = ctypes.c_void_p()
            N.check(N.lib().sacb_get_stream(self._h, ctypes.byref(sp)))
            self._ext_stream = torch.cuda.ExternalStream(sp.value, device=f"cuda:{self._cfg.device}")
        return torch.cuda.stream(self._ext_stream)

    def _tensor(self, net_id, slot, t, shape):
        dev = ctypes.c_void_p()
        N.check(N.lib().sacb_tensor_dev(self._h, 0, net_id, slot, t, ctypes.byref(dev)))
        return torch.as_tensor(N.DevArray(dev.value, tuple(shape), self), device=f"cuda:{self._cfg.device}")

    @property
    def alpha(self):
        """sac_imp.py:23 / :135: the constructor float until the first tuned update, then exp(log_alpha) (1-element tensor)."""
        sc = self._scalars()
        if not self.automatic_entropy_tuning or sc.n_updates == 0 and self._alpha_is_float:
            return float(sc.alpha)
        return torch.tensor([sc.alpha], device=self.device)

    @alpha.setter
    def alpha(self, value):
        sc = self._scalars()
        self._alpha_is_float = not torch.is_tensor(value)
        sc.alpha = float(value.detach().reshape(-1)[0]) if torch.is_tensor(value) else float(value)
        self._set_scalars(sc)

    @property
    def log_alpha(self):
        # a leaf that requires grad, like the reference's (sac_imp.py:48): a checkpoint written here stays trainable when the
        # reference's load_checkpoint (sac_imp.py:222-223) rebinds `self.log_alpha` to the loaded tensor
        return torch.tensor([self._scalars().log_alpha], device=self.device, requires_grad=True)

    @log_alpha.setter
    def log_alpha(self, value):
        sc = self._scalars()
        sc.log_alpha = float(torch.as_tensor(value).detach().reshape(-1)[0])
        self._set_scalars(sc)

    # ---- acting ----------------------------------------------------------------------------------------------
    def select_action(self, state, evaluate=False, *, eps=None):
        """sac_imp.py:54-72: B=1 policy evaluation on the device; `eps` (test hook) replaces the on-device N(0,1) draw."""
        obs = N.f32(state).ravel()
        out = np.empty(self._cfg.act_dim, np.float32)
        e = None if eps is None else N.f32(eps).ravel()
        N.check(N.lib().sacb_select_action(self._h, 0, N.ptr(obs), int(bool(evaluate)), N.ptr(e), N.ptr(out)))
        return out

    # ---- learning --------------------------------------------------------------------------------------------
    def update_parameters(self, batch_size=256, *, eps=None, idx=None, u=None, sync=True):
        """sac_imp.py:74-144.  Draws the minibatch like the reference (`random.sample` positions for the uniform
        buffer, inverse-CDF with `np.random.random_sample` uniforms for the prioritized one), runs the fused update.
        Test hooks: `eps=(eps_next, eps_cur)` [B,act] arrays, `idx` positions, `u` uniforms.  sync=False skips the
        blocking read of the three losses (returns None)."""
        buf = self.replay_buffer
        lib = N.lib()
        self._publish_alias_writes()
        e_next = e_cur = None
        if eps is not None:
            e_next, e_cur = N.f32(eps[0]), N.f32(eps[1])
        losses = np.zeros(3, np.float32)
        flags = 0 if sync else N.NO_LOSS_READBACK
        per = isinstance(buf, PrioritizedReplayBuffer)
        if per:      # host-side draws first: everything behind the flush is then enqueued back to back
            n = min(batch_size, len(buf))
            uu = np.ascontiguousarray(np.random.random_sample(n) if u is None else u, np.float64)
        else:
            ix = buf._draw(batch_size) if idx is None else np.ascontiguousarray(idx, np.int64)
        buf._flush()
        if per:
            N.check(lib.sacb_per_sample(self._h, 0, N.ptr(uu, ctypes.c_double), batch_size, None, None, None, None, None, None, None))
            if self._cfg.per_weighted_loss:      # update_priorities(|q1 - y|): enqueued by the same call, behind the loss copy
                flags |= N.WRITE_BACK_TD
            N.check(lib.sacb_update(self._h, n, None, N.ptr(e_next), N.ptr(e_cur), N.ptr(losses) if sync else None, flags | N.USE_LAST_SAMPLE))
        else:
            N.check(lib.sacb_update(self._h, ix.size, N.ptr(ix, ctypes.c_int64), N.ptr(e_next), N.ptr(e_cur), N.ptr(losses) if sync else None, flags))
        self._alpha_is_float = False
        if not sync:
            return None
        return {"q1_loss": float(losses[0]), "q2_loss": float(losses[1]), "policy_loss": float(losses[2])}

    def learner_steps(self, batch_size=256, k=8, *, sync=True):
        """K learner steps per call, nothing of a step on the host (trainer.py:190-205 with the updates of K
        environment steps batched): prioritized ring = K x the `learner_step` pipeline, uniform ring = K updates whose positions are
        drawn inside the gather stage (a keyed bijection of range(len): distinct positions, like random.sample); eps on the device;
        ONE read-back of the K loss triples.  Bitwise equal to K single calls.  Returns a list of K dicts (sync=False: None)."""
        self.replay_buffer._flush()
        self._publish_alias_writes()
        losses = np.zeros((k, 3), np.float32)
        N.check(N.lib().sacb_update_steps(self._h, batch_size, k, N.ptr(losses) if sync else None, 0 if sync else N.NO_LOSS_READBACK))
        self._alpha_is_float = False
        if not sync:
            return None
        return [{"q1_loss": float(l[0]), "q2_loss": float(l[1]), "policy_loss": float(l[2])} for l in losses]

    def learner_step(self, batch_size=256, *, sync=False):
        """Throughput form of `update_parameters` over the prioritized buffer (per_weighted_loss mode), everything resident
        in HBM: update on the minibatch sampled by the previous call -> priorities <- |q1 - y| -> sample for the next call;
        the write-back and the sample run on a second stream under the rest of the update (`sacb_per_step`).  Device-drawn
        uniforms and eps.  Same values as `update_parameters` with device draws; a transition pushed between two calls
        can first be drawn one call later."""
        buf = self.replay_buffer
        if not isinstance(buf, PrioritizedReplayBuffer):
            raise ValueError("learner_step needs the prioritized replay buffer")
        buf._flush()
        self._publish_alias_writes()
        losses = np.zeros(3, np.float32)
        N.check(N.lib().sacb_per_step(self._h, batch_size, N.ptr(losses) if sync else None, 0 if sync else N.NO_LOSS_READBACK))
        self._alpha_is_float = False
        if not sync:
            return None
        return {"q1_loss": float(losses[0]), "q2_loss": float(losses[1]), "policy_loss": float(losses[2])}

    def update_from_batch(self, batch, eps=None, is_weights=None, export_grads=False, want_td=False):
        """Same step on a caller-supplied minibatch dict(s,a,r,s2,d) (parity tests / benchmarks; no replay involved)."""
        s, a, r, s2, d = (N.f32(batch[k]) for k in ("s", "a", "r", "s2", "d"))
        B = s.shape[0]
        self._publish_alias_writes()
        e_next = e_cur = None
        if eps is not None:
            e_next, e_cur = N.f32(eps[0]), N.f32(eps[1])
        w = None if is_weights is None else N.f32(is_weights)
        losses = np.zeros(3, np.float32)
        td = np.empty(B, np.float32) if want_td else None
        N.check(N.lib().sacb_update_batch(self._h, B, N.ptr(s), N.ptr(a), N.ptr(r), N.ptr(s2), N.ptr(d), N.ptr(w), N.ptr(e_next), N.ptr(e_cur),
                                          N.ptr(losses), N.ptr(td), N.EXPORT_GRADS if export_grads else 0))
        self._alpha_is_float = False
        out = {"q1_loss": float(losses[0]), "q2_loss": float(losses[1]), "policy_loss": float(losses[2])}
        return (out, td) if want_td else out

    def exported_grads(self, net):
        """Gradients of the last update run with export_grads=True, keyed like `named_parameters()`."""
        net_id = {"policy": N.NET_POLICY, "q1": N.NET_Q1, "q2": N.NET_Q2}[net]
        module = getattr(self, net)
        return {name: self._tensor(net_id, N.SLOT_GRAD, i, p.shape).cpu().numpy() for i, (name, p) in enumerate(module.named_parameters())}

    def _soft_update_target_networks(self):
        raise RuntimeError("the Polyak update is fused into update_parameters (sac_imp.py:146-152)")

    # ---- persistence (same dictionary keys as the reference) ----------------------------------------------------
    def save(self, path):
        self.synchronize()                  # an update enqueued with sync=False / learner_step may still be running
        with self._library_stream():
            torch.save({f"{n}_state_dict": getattr(self, n).state_dict() for n in _NETS} | {"alpha": self.alpha}, path)

    def _load_nets(self, checkpoint):
        # the copies into the aliased arena run on the library's own stream (torch's current stream is not ordered against it)
        with self._library_stream():
            for n in _NETS:
                getattr(self, n).load_state_dict(checkpoint[f"{n}_state_dict"])
        self.synchronize()

    def load(self, path):
        checkpoint = torch.load(path, map_location=self.device, weights_only=False)
        self._load_nets(checkpoint)
        self.alpha = checkpoint["alpha"]

    def save_checkpoint(self, path, episode, total_steps, replay_buffer=True, *, rng_state=False):
        """sac_imp.py:177-201, same keys.  rng_state=True adds one extra key ('sacb_rng_state': the Philox counters of the
        on-device eps / exploration draws) so that a seeded run resumes its own random stream; the reference ignores it."""
        self.synchronize()
        checkpoint = {"episode": episode, "total_steps": total_steps}
        with self._library_stream():
            checkpoint.update({f"{n}_state_dict": getattr(self, n).state_dict() for n in _NETS})
            for n in ("policy", "q1", "q2"):
                checkpoint[f"{n}_optimizer_state_dict"] = getattr(self, f"{n}_optimizer").state_dict()
            checkpoint["alpha"] = self.alpha
            if self.automatic_entropy_tuning:
                checkpoint["log_alpha"] = self.log_alpha
                checkpoint["alpha_optimizer_state_dict"] = self.alpha_optimizer.state_dict()
            if replay_buffer:
                checkpoint["replay_buffer"] = self.replay_buffer.buffer
            if rng_state:
                sc = self._scalars()
                checkpoint["sacb_rng_state"] = {"seed": int(self._cfg.seed), "n_updates": int(sc.n_updates), "act_counter": int(sc.act_counter)}
            # the reference only writes the file inside `if replay_buffer:` (sac_imp.py:198-201); always writing is the fix
            torch.save(checkpoint, path)

    def load_checkpoint(self, path, load_replay_buffer=True):
        checkpoint = torch.load(path, map_location=self.device, weights_only=False)   # pickled deque / numpy inside
        self._load_nets(checkpoint)
        for n in ("policy", "q1", "q2"):
            if f"{n}_optimizer_state_dict" in checkpoint:
                getattr(self, f"{n}_optimizer").load_state_dict(checkpoint[f"{n}_optimizer_state_dict"])
        self.alpha = checkpoint["alpha"]
        if self.automatic_entropy_tuning and "log_alpha" in checkpoint:
            self.log_alpha = checkpoint["log_alpha"]
        if "alpha_optimizer_state_dict" in checkpoint and self.automatic_entropy_tuning:
            self.alpha_optimizer.load_state_dict(checkpoint["alpha_optimizer_state_dict"])
        if load_replay_buffer and "replay_buffer" in checkpoint:
            self.replay_buffer.buffer = checkpoint["replay_buffer"]
        rs = checkpoint.get("sacb_rng_state")
        if rs is not None and int(rs.get("seed", -1)) == int(self._cfg.seed):      # resume this seed's Philox streams where they stopped
            sc = self._scalars()
            sc.n_updates, sc.act_counter = int(rs["n_updates"]), int(rs["act_counter"])
            self._set_scalars(sc)
        return checkpoint.get("episode", 0), checkpoint.get("total_steps", 0)

    # ---- instrumentation ------------------------------------------------------------------------------------------
    def stats(self):
        st = N.Stats()
        N.check(N.lib().sacb_get_stats(self._h, ctypes.byref(st)))
        return {k: getattr(st, k) for k, _ in st._fields_}

    def synchronize(self):
        N.check(N.lib().sacb_synchronize(self._h))


class PopulationSAC:
    """`n_agents` independent SAC agents (distinct seeds) trained by ONE batched update program per step on one GPU
    (BASELINE.json configs[4]; the per-agent step is sac_imp.py:74-144).  Agent i = exactly what `torch.manual_seed(seeds[i]);
    SAC(state_dim, action_dim, ...)` builds: same initialiser calls in the same order, its own Adam state, temperature, Philox stream
    (stream id = agent index) and its own uniform replay ring of `capacity` transitions in HBM -- the agent index is a grid
    dimension of every stage kernel and the 4th coordinate of the TMA descriptors.  No inter-agent (or inter-GPU) traffic:
    `distributed.partition_agents` assigns global agent ids to ranks.  Minibatch positions are drawn on the device."""

    def __init__(self, n_agents, state_dim, action_dim, hidden_dim=256, gamma=0.99, tau=0.005, lr=3e-4, alpha=0.2,
                 automatic_entropy_tuning=True, device="cuda", *, seeds=None, capacity=100000, max_batch=256, math="bf16x3",
                 launch="staged", seed=None, action_bounds=None):
        if not str(device).startswith("cuda"):
            raise RuntimeError("PopulationSAC runs on a B200 only (device='cuda[:i]'); there is no CPU path")
        self.n_agents, self.device = int(n_agents), device
        self.seeds = list(range(self.n_agents)) if seeds is None else [int(x) for x in seeds]
        if len(self.seeds) != self.n_agents:
            raise ValueError("one seed per agent")
        kw = {} if action_bounds is None else {"action_bounds": action_bounds}
        probe = GaussianPolicy(state_dim, action_dim, hidden_dim, **kw)
        cfg = N.default_config()
        cfg.obs_dim, cfg.act_dim, cfg.hidden_dim, cfg.n_hidden = state_dim, action_dim, hidden_dim, getattr(probe, "N_HIDDEN", 2)
        cfg.gamma, cfg.tau, cfg.lr, cfg.alpha0, cfg.auto_entropy = gamma, tau, lr, alpha, int(bool(automatic_entropy_tuning))
        cfg.action_scale, cfg.action_bias = probe.action_scale, probe.action_bias
        cfg.replay_kind, cfg.capacity, cfg.max_batch, cfg.n_agents = N.REPLAY_UNIFORM, capacity, max_batch, self.n_agents
        cfg.device = torch.device(device).index or 0
        cfg.math_mode = {"fp32": N.MATH_FP32, "bf16x3": N.MATH_BF16X3}[math]
        cfg.launch_mode = {"staged": N.LAUNCH_STAGED, "persistent": N.LAUNCH_PERSISTENT}[launch]
        cfg.seed = random.getrandbits(63) if seed is None else int(seed)
        self._cfg = cfg
        self._h = N.create(cfg)
        self._names = {}
        lib = N.lib()
        for i, sd in enumerate(self.seeds):
            torch.manual_seed(sd)                      # the construction order of sac_imp.py:28-36 under this agent's seed
            pol = GaussianPolicy(state_dim, action_dim, hidden_dim, **kw)
            q1, q2 = QNetwork(state_dim, action_dim, hidden_dim), QNetwork(state_dim, action_dim, hidden_dim)
            QNetwork(state_dim, action_dim, hidden_dim); QNetwork(state_dim, action_dim, hidden_dim)     # the targets' own (discarded) init draws
            for net_id, mod in ((N.NET_POLICY, pol), (N.NET_Q1, q1), (N.NET_Q2, q2), (N.NET_Q1_TARGET, q1), (N.NET_Q2_TARGET, q2)):
                self._names[net_id] = [(n, tuple(p.shape)) for n, p in mod.named_parameters()]
                for t, (_, p) in enumerate(mod.named_parameters()):
                    host = N.f32(p.detach().numpy())
                    N.check(lib.sacb_import_tensor(self._h, i, net_id, N.SLOT_PARAM, t, N.ptr(host), host.size))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None:
            try:
                N.lib().sacb_destroy(h)
            except Exception:
                pass
            self._h = None

    def __len__(self):
        return self.n_agents

    # ---- replay (per-agent uniform ring, replay_buffer.py:5-22) -------------------------------------------------------
    def push(self, agent, state, action, reward, next_state, done):
        self.push_many(agent, N.f32(state)[None], N.f32(action)[None], [reward], N.f32(next_state)[None], [done])

    def push_many(self, agent, states, actions, rewards, next_states, dones):
        s, a, s2 = N.f32(states), N.f32(actions), N.f32(next_states)
        r, d = N.f32(rewards).ravel(), N.f32(dones).ravel()
        N.check(N.lib().sacb_push(self._h, int(agent), N.ptr(s), N.ptr(a), N.ptr(r), N.ptr(s2), N.ptr(d), s.shape[0]))

    def buffer_len(self, agent):
        return int(N.lib().sacb_len(self._h, int(agent)))

    # ---- acting / learning ------------------------------------------------------------------------------------------------
    def select_action(self, states, evaluate=False, *, eps=None):
        """states [n_agents, obs] -> actions [n_agents, act]: agent i acts on row i with its own policy (sac_imp.py:54-72 each)."""
        obs = N.f32(states).reshape(self.n_agents, self._cfg.obs_dim)
        out = np.empty((self.n_agents, self._cfg.act_dim), np.float32)
        e = None if eps is None else N.f32(eps).reshape(self.n_agents, self._cfg.act_dim)
        N.check(N.lib().sacb_select_action_batch(self._h, N.ptr(obs), int(bool(evaluate)), N.ptr(e), N.ptr(out)))
        return out

    def update_parameters(self, batch_size=256, *, idx=None, eps=None, sync=True):
        """One update of EVERY agent (sac_imp.py:74-144 each) in one program.  idx [n_agents, B] positions / eps = (eps_next,
        eps_cur) [n_agents, B, act] are test hooks; by default both are drawn on the device.  Returns a list of loss dicts."""
        lib = N.lib()
        e_next = e_cur = None
        if eps is not None:
            e_next, e_cur = N.f32(eps[0]), N.f32(eps[1])
        if idx is not None:
            ix = np.ascontiguousarray(idx, np.int64).reshape(self.n_agents, batch_size)
            N.check(lib.sacb_update(self._h, batch_size, N.ptr(ix, ctypes.c_int64), N.ptr(e_next), N.ptr(e_cur), None, N.NO_LOSS_READBACK))
        else:
            N.check(lib.sacb_update(self._h, batch_size, None, N.ptr(e_next), N.ptr(e_cur), None, N.NO_LOSS_READBACK | N.DEVICE_INDICES))
        if not sync:
            return None
        losses = np.zeros((self.n_agents, 3), np.float32)
        N.check(lib.sacb_get_losses_all(self._h, N.ptr(losses)))
        return [{"q1_loss": float(l[0]), "q2_loss": float(l[1]), "policy_loss": float(l[2])} for l in losses]

    def synchronize(self):
        N.check(N.lib().sacb_synchronize(self._h))

    # ---- per-agent state in the reference's save() layout (sac_imp.py:154-162) ---------------------------------------
    def agent_state(self, agent):
        lib, out = N.lib(), {}
        for net_id, name in enumerate(_NETS):
            sd = {}
            for t, (pname, shape) in enumerate(self._names[net_id]):
                host = np.empty(shape, np.float32)
                N.check(lib.sacb_export_tensor(self._h, int(agent), net_id, N.SLOT_PARAM, t, N.ptr(host), host.size))
                sd[pname] = torch.from_numpy(host)
            out[f"{name}_state_dict"] = sd
        sc = N.Scalars()
        N.check(lib.sacb_get_scalars(self._h, int(agent), ctypes.byref(sc)))
        out["alpha"] = float(sc.alpha) if sc.n_updates == 0 or not self._cfg.auto_entropy else torch.tensor([sc.alpha])
        return out

    def load_agent(self, agent, checkpoint):
        """Inverse of `agent_state` (accepts a file written by the reference's or the product's `SAC.save`)."""
        lib = N.lib()
        for net_id, name in enumerate(_NETS):
            sd = checkpoint[f"{name}_state_dict"]
            for t, (pname, _) in enumerate(self._names[net_id]):
                host = N.f32(torch.as_tensor(sd[pname]).detach().cpu().numpy())
                N.check(lib.sacb_import_tensor(self._h, int(agent), net_id, N.SLOT_PARAM, t, N.ptr(host), host.size))
        sc = N.Scalars()
        N.check(lib.sacb_get_scalars(self._h, int(agent), ctypes.byref(sc)))
        a = checkpoint["alpha"]
        sc.alpha = float(a.detach().reshape(-1)[0]) if torch.is_tensor(a) else float(a)
        N.check(lib.sacb_set_scalars(self._h, int(agent), ctypes.byref(sc)))
