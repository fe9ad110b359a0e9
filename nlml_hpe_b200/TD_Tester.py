"""Drop-in for the reference's TD_Tester module (the Tucker-fit entry points), B200-backed.

Same names, argument meaning and return conventions as /root/reference/TD_Tester.py:
  optimize_with_sgd(W, x, u_id, u_id_shape, params_y, params_p, params_r,
                    learning_rate=0.001, num_iterations=3000) -> torch.float32[3+u_id_shape]   (:127-159)
  Test(W, x, u_id_shape, optimized_params_y, optimized_params_p, optimized_params_r,
       u_id, f_y, f_p, f_r) -> (yaw_deg, pitch_deg, roll_deg, u_id)                           (:162-291)
plus the batched form the reference lacks (`optimize_with_sgd_batch`).

Behavioural notes:
* `Test` returns what the reference's `Test` returns, BIT FOR BIT: the default `TEST_SOLVER = "powell"` runs scipy's modified
  Powell search (the reference's minimize(..., method='Powell'), :191-194) on the GPU -- algorithm restated from scipy
  1.18.1, objective evaluated in the reference's own float64 operation order (csrc/powell_math.h) -- pinned by 96 outputs of
  the real reference including the evaluation counts (tests/golden/powell_golden.npz).  Two faster alternatives:
  `TEST_SOLVER = "converged"`: damped-Newton solve of the same objective (16 M poses/s batched).  It does NOT return
  Powell's angles: the objective value at its result is never above Powell's (same basin, < 1 % apart), but Powell stops
  at xtol = ftol = 1e-4 in a flat valley, so the angles differ by 0.28 degrees in the median, 2.8 at the 90th percentile,
  8.7 at most over those 96 samples; a warning is raised if the solve hits its evaluation cap.
  `TEST_SOLVER = "sgd"`: the fixed-iteration block the reference keeps commented out at :168-184 (<= 1e-2 degrees).
* Like the reference, `Test` returns the u_id it was GIVEN (normally None), not the optimum (:291).
* No module-global trace lists (:18-22) are kept: the call is re-entrant.
* No progress printing (:131, :142-143).
"""
from __future__ import annotations

import hashlib
import warnings

import numpy as np
import torch

from .tucker import TuckerFitter

TEST_SOLVER = "powell"      # "powell" (the reference's shipped default, bit for bit), "converged" or "sgd": what Test() runs
SOLVE_MAX_EVALS = 64        # evaluation cap of the converged solve (csrc/tucker_math.h lm_default_options)

_PLAN_CACHE = {}
_PLAN_CACHE_MAX = 4


def _as_numpy(a, dtype):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=dtype))


_FAST = {}   # (id(W), data pointer, shape) -> (cheap digest, full key): skips re-hashing an array object already seen


def _cheap_digest(W):
    """Three vector reductions over W's words (0.1 ms for the 758 KB shipped tensor): detects an in-place change of an
    array object that was hashed before.  It is NOT the cache key (the key is the cryptographic digest below)."""
    w = W.reshape(-1).view(np.uint32)
    return (int(np.add.reduce(w, dtype=np.uint64)), int(np.bitwise_xor.reduce(w)), int(np.add.reduce(w[::7], dtype=np.uint64)))


def _content_key(W, rows, device):
    """Cache key of one set of constants: BLAKE2b over the full contents of W (float32, contiguous) and the cosine rows
    (float64) -- two different tensors cannot share a plan.  The reference passes W on every call (TD_Inference.py:56),
    so the 2 ms hash is taken once per array OBJECT: a later call with the same object (same id, buffer address and
    shape, unchanged cheap digest) reuses its key."""
    ident = (id(W), W.ctypes.data, W.shape)
    cheap = _cheap_digest(W)
    hit = _FAST.get(ident)
    if hit is not None and hit[0] == cheap:
        wkey = hit[1]
    else:
        wkey = hashlib.blake2b(W.tobytes(), digest_size=20).hexdigest()
        if len(_FAST) > 16:
            _FAST.clear()
        _FAST[ident] = (cheap, wkey)
    h = hashlib.blake2b(digest_size=20)
    h.update(str(W.shape).encode())
    h.update(wkey.encode())
    for r in rows:
        h.update(str(r.shape).encode())
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
    if TEST_SOLVER == "powell":
        fit = _fitter(W, optimized_params_y, optimized_params_p, optimized_params_r)
        if fit.ranks[0] != int(u_id_shape):
            raise ValueError(f"u_id_shape={u_id_shape} does not match W.shape[0]={fit.ranks[0]}")
        xg = torch.from_numpy(_as_numpy(x, np.float32).reshape(1, -1)).to(fit.device)
        deg = np.degrees(fit.powell(xg)[0].cpu().numpy())          # np.degrees(result.x), :196
        return deg[0], deg[1], deg[2], u_id
    if TEST_SOLVER == "sgd":
        p = optimize_with_sgd(W, x, u_id, u_id_shape, optimized_params_y, optimized_params_p, optimized_params_r).numpy()
    elif TEST_SOLVER == "converged":
        fit = _fitter(W, optimized_params_y, optimized_params_p, optimized_params_r)
        if fit.ranks[0] != int(u_id_shape):
            raise ValueError(f"u_id_shape={u_id_shape} does not match W.shape[0]={fit.ranks[0]}")
        xg = torch.from_numpy(_as_numpy(x, np.float32).reshape(1, -1)).to(fit.device)
        P, evals = fit.solve(xg, return_evals=True)
        if int(evals[0].item()) >= SOLVE_MAX_EVALS:
            warnings.warn(f"TD_Tester.Test: the converged fit used all {SOLVE_MAX_EVALS} evaluations without meeting its "
                          "step tolerance; the returned angles are the last iterate", RuntimeWarning)
        p = P[0].cpu().numpy()
    else:
        raise ValueError(f"TEST_SOLVER must be 'powell', 'converged' or 'sgd', got {TEST_SOLVER!r}")
    deg = np.degrees(p.astype(np.float64))
    return deg[0], deg[1], deg[2], u_id
