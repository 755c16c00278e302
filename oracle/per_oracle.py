"""ctypes loader for oracle/per_oracle.c + the deque/ring bookkeeping of the reference buffers.
TEST INFRASTRUCTURE ONLY (see per_oracle.c header).  Parity status: PINNED (tests/golden/per_*.npz)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libper_oracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "per_oracle.c")):
            build()
        L = ctypes.CDLL(path)
        f32p, f64p, i64p = (ctypes.POINTER(t) for t in (ctypes.c_float, ctypes.c_double, ctypes.c_int64))
        L.per_oracle_pairwise_sum_f32.restype = ctypes.c_float
        L.per_oracle_pairwise_sum_f32.argtypes = [f32p, ctypes.c_int64]
        L.per_oracle_beta.restype = ctypes.c_double
        L.per_oracle_beta.argtypes = [ctypes.c_double, ctypes.c_int64, ctypes.c_int64]
        L.per_oracle_sample.restype = ctypes.c_int
        L.per_oracle_sample.argtypes = [f32p, ctypes.c_int64, f64p, ctypes.c_int64, ctypes.c_double, i64p, f32p, f32p, f64p]
        L.per_oracle_update_priorities.restype = None
        L.per_oracle_update_priorities.argtypes = [f32p, i64p, f32p, ctypes.c_int64]
        L.per_oracle_push.restype = ctypes.c_int64
        L.per_oracle_push.argtypes = [f32p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64]
        L.per_oracle_pow_alpha.restype = None
        L.per_oracle_pow_alpha.argtypes = [f32p, ctypes.c_int64, ctypes.c_double, f32p]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def pairwise_sum(a):
    a = np.ascontiguousarray(a, np.float32)
    return np.float32(lib().per_oracle_pairwise_sum_f32(_p(a, ctypes.c_float), a.size))


def beta(frame, beta_start=0.4, beta_frames=100000):
    return lib().per_oracle_beta(beta_start, beta_frames, frame)


def sample(p_alpha, u, beta_value, want_tables=False):
    """replay_buffer.py:57-68 downstream of the p**alpha table -> (idx int64[k], weights float32[k])."""
    p_alpha = np.ascontiguousarray(p_alpha, np.float32)
    u = np.ascontiguousarray(u, np.float64)
    n, k = p_alpha.size, u.size
    idx = np.empty(k, np.int64)
    w = np.empty(k, np.float32)
    probs = np.empty(n, np.float32) if want_tables else None
    cdf = np.empty(n, np.float64) if want_tables else None
    rc = lib().per_oracle_sample(_p(p_alpha, ctypes.c_float), n, _p(u, ctypes.c_double), k, float(beta_value),
                                 _p(idx, ctypes.c_int64), _p(w, ctypes.c_float),
                                 _p(probs, ctypes.c_float) if want_tables else None,
                                 _p(cdf, ctypes.c_double) if want_tables else None)
    assert rc == 0
    return (idx, w, probs, cdf) if want_tables else (idx, w)


def update_priorities(priorities, idx, td):
    idx = np.ascontiguousarray(idx, np.int64)
    td = np.ascontiguousarray(td, np.float32)
    assert priorities.dtype == np.float32 and priorities.flags.c_contiguous
    lib().per_oracle_update_priorities(_p(priorities, ctypes.c_float), _p(idx, ctypes.c_int64), _p(td, ctypes.c_float), idx.size)


def push(priorities, length, pos):
    """Returns the new pos; caller bumps length (replay_buffer.py:36-46)."""
    return lib().per_oracle_push(_p(priorities, ctypes.c_float), priorities.size, length, pos)


def pow_alpha(priorities, alpha=0.6):
    priorities = np.ascontiguousarray(priorities, np.float32)
    out = np.empty_like(priorities)
    lib().per_oracle_pow_alpha(_p(priorities, ctypes.c_float), priorities.size, alpha, _p(out, ctypes.c_float))
    return out


def uniform_draws(seed, k):
    """The float64 uniforms np.random.choice consumes after np.random.seed(seed) (mtrand.pyx::choice)."""
    return np.random.RandomState(seed).random_sample(k)


class DequeRing:
    """Index algebra of `deque(maxlen=capacity)` (replay_buffer.py:7-11): logical j = j-th oldest."""

    def __init__(self, capacity):
        self.capacity, self.count, self.head = capacity, 0, 0   # head = physical slot of the oldest entry

    def push_slot(self):
        if self.count < self.capacity:
            slot = (self.head + self.count) % self.capacity
            self.count += 1
        else:
            slot = self.head
            self.head = (self.head + 1) % self.capacity
        return slot

    def physical(self, j):
        return (self.head + np.asarray(j)) % self.capacity
