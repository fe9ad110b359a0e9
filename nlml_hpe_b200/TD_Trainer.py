"""Drop-in for the reference's TD_Trainer module (cosine-curve fit of the Tucker factor columns), B200-backed.

Reference (/root/reference/TD_Trainer.py): Train(yaw_params, pitch_params, roll_params) (:232-351) takes three
(U, w) pairs -- U [n_bins, rank] a factor matrix of the Tucker decomposition, w the angle bins in degrees -- and fits
a cos(b w + c) + d to every column: Fourier initial guess (est_params_by_Uniform_Fourier, :125-148), then scipy Powell on
the least-squares objective (:38-43, :60-93).  It runs once, offline, and produces the optimized_{yaw,pitch,roll} rows
the inference hot path loads (TD_main.py:254, :262).  Here both steps run on the GPU in float64 (one thread per column,
csrc/powell_math.h), reproducing the reference's shipped rows to 1e-6 (tests/test_trainer.py).
compute_W is W = core x_5 U_feat (TD_main.py:231-238).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .tucker import _device_index


def _fit(U, w, device=None, return_info=False):
    lib = _lib.load()
    U = np.ascontiguousarray(np.asarray(U, dtype=np.float64))
    w = np.ascontiguousarray(np.asarray(w, dtype=np.float64))
    if U.ndim != 2 or w.ndim != 1 or w.shape[0] != U.shape[0]:
        raise ValueError(f"expected U [n_bins, rank] and w [n_bins], got {U.shape} and {w.shape}")
    n_rows, n_cols = U.shape
    out, init = np.empty((n_cols, 4)), np.empty((n_cols, 4))
    fun, nfev = np.empty(n_cols), np.empty(n_cols, dtype=np.int32)
    _lib.check(lib.nlml_cosine_fit_f64(U.ctypes.data, n_rows, n_cols, w.ctypes.data, out.ctypes.data, init.ctypes.data,
                                       fun.ctypes.data, nfev.ctypes.data, _device_index(device)))
    return (out, init, fun, nfev) if return_info else out


def est_params_by_Uniform_Fourier(U, w, device=None):
    """Initial (a, b, c, d) per column from the dominant non-zero DFT frequency (:125-148)."""
    return _fit(U, w, device, return_info=True)[1]


def estimate_init_Fourier_Trans(yaw_params, pitch_params, roll_params, device=None):
    """(:150-160)"""
    return tuple(est_params_by_Uniform_Fourier(U, w, device) for U, w in (yaw_params, pitch_params, roll_params))


def optimize_for_matrix_using_grads(U_matrix, w_vector, initial_guesses=None, device=None):
    """(:60-93).  The initial guesses are recomputed on the device (they are a function of U and w, :241)."""
    return _fit(U_matrix, w_vector, device)


def Train(yaw_params, pitch_params, roll_params, device=None):
    """Reference entry point (:232): -> (optimized_params_yaw, optimized_params_pitch, optimized_params_roll),
    each [rank, 4] float64 rows (a, b, c, d).  Prints the three blocks like the reference (:316-323)."""
    res = tuple(_fit(U, w, device) for U, w in (yaw_params, pitch_params, roll_params))
    print("Optimal parameters for yaw:")
    print(res[0])
    print("\nOptimal parameters for pitch:")
    print(res[1])
    print("\nOptimal parameters for roll:")
    print(res[2])
    return res


def compute_W(core_tensor, feature_matrix, device=None):
    """W = core x_5 U_feat (TD_main.py:231-238: tl.tensordot(core, U_feat^T, axes=(4, 0))).
    core_tensor float32 [R_id, R_y, R_p, R_r, R_feat], feature_matrix float32 [F, R_feat] -> W [R_id, R_y, R_p, R_r, F]."""
    lib = _lib.load()
    core = np.ascontiguousarray(np.asarray(core_tensor, dtype=np.float32))
    Uf = np.ascontiguousarray(np.asarray(feature_matrix, dtype=np.float32))
    if core.ndim != 5 or Uf.ndim != 2 or Uf.shape[1] != core.shape[4]:
        raise ValueError(f"expected core [..., R_feat] and feature matrix [F, R_feat], got {core.shape} and {Uf.shape}")
    R, M, F = int(np.prod(core.shape[:4])), core.shape[4], Uf.shape[0]
    W = np.empty((R, F), dtype=np.float32)
    _lib.check(lib.nlml_core_times_features_f32(core.ctypes.data, Uf.ctypes.data, R, M, F, W.ctypes.data, _device_index(device)))
    return W.reshape(*core.shape[:4], F)
