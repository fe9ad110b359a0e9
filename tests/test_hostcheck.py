"""The kernels' per-sample arithmetic (csrc/tucker_math.h), compiled for the host, against the
reference goldens.  Same statements the CUDA kernels execute; only libm's sinf/cosf differ."""
import ctypes

import numpy as np

DEG = 180.0 / np.pi


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _fit(lib, W, rows, X, T):
    W2 = np.ascontiguousarray(W.reshape(-1, W.shape[-1]))
    X = np.ascontiguousarray(X)
    P = np.zeros((len(X), 8), np.float32)
    lib.hostcheck_tucker_fit_5333(_ptr(W2), W2.shape[1], _ptr(rows[0]), _ptr(rows[1]), _ptr(rows[2]), _ptr(X),
                                  ctypes.c_int64(len(X)), ctypes.c_int64(X.shape[1]), T, ctypes.c_float(1e-3),
                                  ctypes.c_float(1.0), _ptr(P))
    return P


def test_folded_gram_gradient_matches_autograd(hostcheck, art, rows, X1k, tucker_golden):
    W2 = np.ascontiguousarray(art["W"].reshape(135, 1404))
    Pq = np.ascontiguousarray(tucker_golden["grad_P"])
    G = np.zeros_like(Pq)
    X = np.ascontiguousarray(X1k[:32])
    hostcheck.hostcheck_tucker_grad_5333(_ptr(W2), 1404, _ptr(rows[0]), _ptr(rows[1]), _ptr(rows[2]), _ptr(X),
                                         ctypes.c_int64(32), ctypes.c_int64(1404), _ptr(Pq), _ptr(G))
    ref = tucker_golden["grad_G"]
    assert np.abs(G - ref).max() / np.abs(ref).max() < 1e-5


def test_fit_matches_reference_sgd3000_shipped(hostcheck, art, rows, X1k, tucker_golden):
    P = _fit(hostcheck, art["W"], rows, X1k[:16], 3000)
    ref = tucker_golden["sgd3000_shipped_P"]
    assert np.abs(P[:, :3] - ref[:, :3]).max() * DEG < 1e-3     # 10x inside the 1e-2 deg tolerance
    assert np.abs(P[:, 3:] - ref[:, 3:]).max() < 1e-5


def test_fit_matches_reference_sgd200(hostcheck, art, rows, X1k, tucker_golden):
    idx = tucker_golden["sgd200_shipped_idx"]
    P = _fit(hostcheck, art["W"], rows, X1k[idx], 200)
    ref = tucker_golden["sgd200_shipped_P"]
    d = np.abs(P[:, :3] - ref[:, :3]).max(1) * DEG   # transiently ill-conditioned samples: see test_oracle.py
    assert np.quantile(d, 0.9) < 1e-3 and d.max() < 5e-2


def test_fit_matches_reference_synthetic_core(hostcheck, rows, tucker_golden):
    from nlml_hpe_b200 import synthetic
    G = synthetic.synthetic_core((5, 3, 3, 3), 1404, seed=7)
    Xg = synthetic.make_features(1000, G, *rows, U_id=None, seed=4321)
    P = _fit(hostcheck, G, rows, Xg[:8], 3000)
    ref = tucker_golden["sgd3000_syncore_P"]
    assert np.abs(P[:, :3] - ref[:, :3]).max() * DEG < 1e-2


def test_fit_edge_inputs(hostcheck, art, rows, tucker_golden):
    """noise-free, all-zero ('no face' sentinel) and 10x-scaled inputs."""
    P = _fit(hostcheck, art["W"], rows, tucker_golden["sgd500_edge_X"], 500)
    ref = tucker_golden["sgd500_edge_P"]
    assert np.abs(P[:, :3] - ref[:, :3]).max() * DEG < 1e-2
    assert np.abs(P[1]).max() == 0.0 and np.abs(ref[1]).max() == 0.0   # zero input never moves


def test_sincos_small_accuracy(hostcheck):
    """The tensor-core kernel's slow-path-free sincos: <= 2 ulp-ish absolute error on the argument range of the
    cosine factors (b*w + c stays within a few radians; checked out to +-100)."""
    import ctypes
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-8, 8, 200000), rng.uniform(-100, 100, 50000),
                        np.array([0.0, -0.0, np.pi / 4, -np.pi / 4, np.pi / 2, np.pi, 1e-8, -1e-8])]).astype(np.float32)
    sn, cs = np.empty_like(x), np.empty_like(x)
    hostcheck.hostcheck_sincos_small.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    hostcheck.hostcheck_sincos_small.restype = None
    hostcheck.hostcheck_sincos_small(x.ctypes.data, x.size, sn.ctypes.data, cs.ctypes.data)
    xd = x.astype(np.float64)
    assert np.abs(sn - np.sin(xd)).max() < 2.5e-7
    assert np.abs(cs - np.cos(xd)).max() < 2.5e-7
