"""Drop-in for the reference's TD_Tester module (the Tucker-fit entry points), B200-backed.

Same names, argument meaning and return conventions as /root/reference/TD_Tester.py:
  optimize_with_sgd(W, x, u_id, u_id_shape, params_y, params_p, params_r,
                    learning_rate=0.001, num_iterations=3000) -> torch.float32[3+u_id_shape]   (:127-159)
  Test(W, x, u_id_shape, optimized_params_y, optimized_params_p, optimized_params_r,
       u_id, f_y, f_p, f_r) -> (yaw_deg, pitch_deg, roll_deg, u_id)                           (:162-291)
plus the batched form the reference lacks (`optimize_with_sgd_batch`).

Behavioural notes:
* `Test` runs, like the reference's default, a CONVERGED fit from p = 0: where the reference calls scipy Powell
  (:191-199; unpinned third-party search, stops at xtol = ftol = 1e-4) this module runs the damped-Newton solve
  of the same objective on the GPU (`solve_batch`), which ends in the same basin at a lower or equal loss
  (tests/test_tucker_gpu.py).  `TEST_SOLVER = "sgd"` selects the fixed-iteration block the reference keeps
  commented out at :168-184 instead (the one with the bit-level parity contract).
* Like the reference, `Test` returns the u_id it was GIVEN (normally None), not the optimum (:291).
* No module-global trace lists (:18-22) are kept: the call is re-entrant.
* No progress printing (:131, :142-143).
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch

from .tucker import TuckerFitter

TEST_SOLVER = "converged"   # or "sgd": what Test() runs (see the module docstring)

_PLAN_CACHE = {}
_PLAN_CACHE_MAX = 4


def _as_numpy(a, dtype):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=dtype))


_WEIGHTS = {}


def _position_weights(n):
    w = _WEIGHTS.get(n)
    if w is None:
        if len(_WEIGHTS) > 8:
            _WEIGHTS.clear()
        w = _WEIGHTS[n] = (np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) | np.uint64(1))
    return w


def _content_key(W, rows, device):
    """Cache key of one set of constants (W float32 contiguous, rows float64): shape + cheap content digests.
    The reference passes W on every call (TD_Inference.py:56), so this runs per sample: three vector reductions
    over the 758 KB (0.1 ms) instead of a cryptographic hash of it (2 ms)."""
    w64 = W.reshape(-1).view(np.uint32).astype(np.uint64, copy=False) if W.size % 2 else W.reshape(-1).view(np.uint64)
    h = hashlib.blake2b(digest_size=16)
    h.update(str(W.shape).encode())
    weights = _position_weights(w64.size)   # odd multipliers: a permutation of the words changes the weighted sum
    h.update(np.array([np.add.reduce(w64), np.bitwise_xor.reduce(w64), np.add.reduce(w64 * weights)], dtype=np.uint64).tobytes())
    for r in rows:
        h.update(r.tobytes())
    return (h.hexdigest(), str(device))


def _fitter(W, params_y, params_p, params_r, device=None):
    W = _as_numpy(W, np.float32)
    rows = [_as_numpy(p, np.float64) for p in (params_y, params_p, params_r)]
    key = _content_key(W, rows, device)
    fit = _PLAN_CACHE.get(key)
    if fit is None:
        if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE))).close()
        fit = _PLAN_CACHE[key] = TuckerFitter(W, *rows, device=device)
    return fit


def optimize_with_sgd_batch(W, X, u_id_shape, params_y, params_p, params_r, learning_rate=0.001,
                            num_iterations=3000, max_norm=1.0, device=None):
    """Batched optimize_with_sgd: X [N,F] (CUDA tensor stays on device; numpy / CPU goes through the
    pipelined host path).  Returns [N, 3+u_id_shape] of the same kind as X."""
    fit = _fitter(W, params_y, params_p, params_r, device)
    if fit.ranks[0] != int(u_id_shape):
        raise ValueError(f"u_id_shape={u_id_shape} does not match W.shape[0]={fit.ranks[0]}")
    if isinstance(X, torch.Tensor) and X.is_cuda:
        return fit.fit(X, num_iterations, learning_rate, max_norm)
    was_tensor = isinstance(X, torch.Tensor)
    out = fit.fit_host(_as_numpy(X, np.float32), num_iterations, learning_rate, max_norm)
    return torch.from_numpy(out) if was_tensor else out


def solve_batch(W, X, u_id_shape, params_y, params_p, params_r, max_evals=0, device=None):
    """Batched converged fit (the optimum Test() looks for, TD_Tester.py:191-199): X [N,F] -> [N, 3+u_id_shape]
    radians + identity coefficients, of the same kind as X."""
    fit = _fitter(W, params_y, params_p, params_r, device)
    if fit.ranks[0] != int(u_id_shape):
        raise ValueError(f"u_id_shape={u_id_shape} does not match W.shape[0]={fit.ranks[0]}")
    if isinstance(X, torch.Tensor) and X.is_cuda:
        return fit.solve(X, max_evals)
    was_tensor = isinstance(X, torch.Tensor)
    out = fit.solve_host(_as_numpy(X, np.float32), max_evals)
    return torch.from_numpy(out) if was_tensor else out


def optimize_with_sgd(W, x, u_id, u_id_shape, params_y, params_p, params_r, learning_rate=0.001,
                      num_iterations=3000):
    """Single-sample form with the reference signature (TD_Tester.py:127).  `u_id` is unused there too."""
    x = _as_numpy(x, np.float32).reshape(1, -1)
    P = optimize_with_sgd_batch(W, x, u_id_shape, params_y, params_p, params_r, learning_rate, num_iterations)
    return torch.from_numpy(np.asarray(P)[0].copy())


def Test(W, x, u_id_shape, optimized_params_y, optimized_params_p, optimized_params_r, u_id, f_y, f_p, f_r):
    """Reference entry point (TD_Tester.py:162-163) -> (yaw, pitch, roll) in degrees and the given u_id."""
    if TEST_SOLVER == "sgd":
        p = optimize_with_sgd(W, x, u_id, u_id_shape, optimized_params_y, optimized_params_p, optimized_params_r).numpy()
    elif TEST_SOLVER == "converged":
        p = np.asarray(solve_batch(W, _as_numpy(x, np.float32).reshape(1, -1), u_id_shape, optimized_params_y,
                                   optimized_params_p, optimized_params_r))[0]
    else:
        raise ValueError(f"TEST_SOLVER must be 'converged' or 'sgd', got {TEST_SOLVER!r}")
    deg = np.degrees(p.astype(np.float64))
    return deg[0], deg[1], deg[2], u_id
