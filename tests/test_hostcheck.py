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


# ---- converged fit (tucker_newton_eval / tucker_lm_solve), host build of the kernel statements ----

def _unpack(H):
    out = np.zeros((len(H), 8, 8))
    for r in range(8):
        for c in range(r + 1):
            out[:, r, c] = out[:, c, r] = H[:, r * (r + 1) // 2 + c]
    return out


def test_newton_terms_match_autograd(hostcheck, art, rows, X1k, tucker_golden):
    """Value, gradient and exact Hessian from ONE pass over the folded Gram tensor vs torch.func on the reference
    objective (TD_Tester.py:110-125) in float64."""
    from oracle import tucker_oracle
    n = 32
    W2 = np.ascontiguousarray(art["W"].reshape(135, 1404))
    Pq = np.ascontiguousarray(tucker_golden["grad_P"][:n])
    X = np.ascontiguousarray(X1k[:n])
    L, G, H = np.zeros(n, np.float32), np.zeros((n, 8), np.float32), np.zeros((n, 36), np.float32)
    hostcheck.hostcheck_tucker_newton_5333(_ptr(W2), 1404, _ptr(rows[0]), _ptr(rows[1]), _ptr(rows[2]), _ptr(X),
                                           ctypes.c_int64(n), ctypes.c_int64(1404), _ptr(Pq), _ptr(L), _ptr(G), _ptr(H))
    Lr, Gr, Hr = tucker_oracle.newton_terms(Pq, art["W"], X, *rows)
    assert np.abs(L - (Lr - 0.5 * (X.astype(np.float64) ** 2).sum(1))).max() < 1e-4
    assert np.abs(G - Gr).max() / np.abs(Gr).max() < 1e-5
    Hf = _unpack(H)
    for blk in ((slice(0, 3), slice(0, 3)), (slice(3, 8), slice(0, 3)), (slice(3, 8), slice(3, 8))):
        assert np.abs(Hf[:, blk[0], blk[1]] - Hr[:, blk[0], blk[1]]).max() / np.abs(Hr[:, blk[0], blk[1]]).max() < 1e-5


def _solve(lib, W, rows, X):
    W2 = np.ascontiguousarray(W.reshape(-1, W.shape[-1]))
    X = np.ascontiguousarray(X)
    P, ev = np.zeros((len(X), 8), np.float32), np.zeros(len(X), np.int32)
    lib.hostcheck_tucker_solve_5333(_ptr(W2), W2.shape[1], _ptr(rows[0]), _ptr(rows[1]), _ptr(rows[2]), _ptr(X),
                                    ctypes.c_int64(len(X)), ctypes.c_int64(X.shape[1]), 0, _ptr(P), _ptr(ev), None, None)
    return P, ev


def test_solve_matches_f64_oracle_and_powell(hostcheck, art, rows, X1k, tucker_golden):
    from oracle import tucker_oracle
    n = 200
    P, ev = _solve(hostcheck, art["W"], rows, X1k[:n])
    ref, _, ev64 = tucker_oracle.lm_fit(art["W"], X1k[:n], *rows)
    d = np.abs(P[:, :3] - ref[:, :3]).max(1) * DEG
    # FP32 floor of this objective: its valley is so flat (a 2 degree move changes L by ~1e-5) that on the worst-
    # conditioned 1% of samples the FP32 gradient noise (~1e-6) moves the optimum by ~1e-2 degrees
    assert d.max() < 5e-2 and np.quantile(d, 0.95) < 1e-2 and np.median(d) < 1e-3
    assert ev.max() <= 64 and ev.mean() < 25
    # same basin as the reference's scipy Powell (Test(), TD_Tester.py:191-199); Powell stops early (ftol 1e-4)
    assert np.abs(P[:8, :3] * DEG - tucker_golden["powell_shipped_deg"]).max() < 5.0


def test_solve_synthetic_core(hostcheck, rows):
    """BASELINE.json config 2 core (random, well conditioned): the FP32 solve and the float64 restatement agree far
    below the budget -- the 1e-2 degree floor on the shipped W is that core's conditioning, not the solver."""
    from nlml_hpe_b200 import synthetic
    from oracle import tucker_oracle
    G = synthetic.synthetic_core((5, 3, 3, 3), 1404, seed=7)
    Xg = synthetic.make_features(128, G, *rows, U_id=None, seed=4321)
    P, ev = _solve(hostcheck, G, rows, Xg)
    ref, _, _ = tucker_oracle.lm_fit(G, Xg, *rows)
    d = np.abs(P[:, :3] - ref[:, :3]).max(1) * DEG
    assert d.max() < 2e-3 and np.median(d) < 2e-4


def test_packed_cholesky_solve(hostcheck):
    """The 8x8 register Cholesky of the converged solver against numpy, and its pivot test on indefinite matrices."""
    rng = np.random.default_rng(11)
    hostcheck.hostcheck_chol_solve8.restype = ctypes.c_int
    for trial in range(50):
        B = rng.standard_normal((8, 8))
        A = B @ B.T + 0.5 * np.eye(8)
        A *= 10.0 ** rng.integers(-3, 4)
        g = rng.standard_normal(8)
        packed = np.array([A[r, c] for r in range(8) for c in range(r + 1)], np.float32)
        d = np.zeros(8, np.float32)
        ok = hostcheck.hostcheck_chol_solve8(_ptr(packed), _ptr(g.astype(np.float32)), _ptr(d))
        ref = np.linalg.solve(A, g)
        assert ok == 1 and np.abs(d - ref).max() <= 2e-4 * np.abs(ref).max() * np.linalg.cond(A) ** 0.5
    A = np.diag([1.0, 2.0, -0.5, 1.0, 1.0, 1.0, 1.0, 1.0])           # one negative pivot -> the solver raises its damping
    packed = np.array([A[r, c] for r in range(8) for c in range(r + 1)], np.float32)
    assert hostcheck.hostcheck_chol_solve8(_ptr(packed), _ptr(np.ones(8, np.float32)), _ptr(np.zeros(8, np.float32))) == 0


def test_folded_form_on_random_cores_and_feature_counts(hostcheck):
    """Value / gradient / Hessian from the folded Gram tensor against torch.func on the reference objective for random
    cores with other feature counts (F = 5, 24, 97) and random cosine rows: nothing in the algebra depends on the shipped
    artefacts."""
    from oracle import tucker_oracle
    rng = np.random.default_rng(2024)
    for F in (5, 24, 97):
        W = rng.standard_normal((5, 3, 3, 3, F)).astype(np.float32)
        rws = [np.ascontiguousarray(np.stack([rng.uniform(-1, 1, 3), rng.uniform(0.5, 3, 3), rng.uniform(-np.pi, np.pi, 3),
                                              rng.uniform(-0.2, 0.2, 3)], 1)) for _ in range(3)]
        n = 16
        X = np.ascontiguousarray(rng.standard_normal((n, F)).astype(np.float32))
        Pq = np.ascontiguousarray(np.concatenate([rng.uniform(-1, 1, (n, 3)), rng.normal(0, 0.3, (n, 5))], 1).astype(np.float32))
        W2 = np.ascontiguousarray(W.reshape(135, F))
        L, G, H = np.zeros(n, np.float32), np.zeros((n, 8), np.float32), np.zeros((n, 36), np.float32)
        hostcheck.hostcheck_tucker_newton_5333(_ptr(W2), F, _ptr(rws[0]), _ptr(rws[1]), _ptr(rws[2]), _ptr(X),
                                               ctypes.c_int64(n), ctypes.c_int64(F), _ptr(Pq), _ptr(L), _ptr(G), _ptr(H))
        Lr, Gr, Hr = tucker_oracle.newton_terms(Pq, W, X, *rws)
        scale = np.abs(Lr).max()
        assert np.abs(L - (Lr - 0.5 * (X.astype(np.float64) ** 2).sum(1))).max() < 2e-5 * scale
        assert np.abs(G - Gr).max() < 2e-5 * np.abs(Gr).max()
        assert np.abs(_unpack(H) - Hr).max() < 2e-5 * np.abs(Hr).max()


def test_solve_vs_reference_powell_96_host(hostcheck, art, rows, X1k, powell_golden):
    """CPU tier of tests/test_tucker_gpu.py::test_solve_vs_reference_powell_96: the converged solver (host build of the
    kernel's statements) against 96 outputs of the reference's scipy-Powell Test(): never a higher objective value,
    angle gap distribution as documented (median 0.28, 90 % 2.8, max 8.7 degrees)."""
    from oracle import tucker_oracle
    idx = powell_golden["idx"]
    X = X1k[idx]
    res = _solve(hostcheck, art["W"], rows, X)
    P = np.asarray(res[0] if isinstance(res, tuple) else res)
    L_ours = tucker_oracle.newton_terms(P, art["W"], X, *rows)[0]
    L_powell = tucker_oracle.newton_terms(powell_golden["p"], art["W"], X, *rows)[0]
    assert np.abs(L_powell - powell_golden["loss"]).max() < 1e-8
    assert (L_ours <= L_powell + 1e-7).all()
    gap = np.abs(np.degrees(P[:, :3].astype(np.float64)) - powell_golden["deg"]).max(1)
    assert np.median(gap) < 0.6 and np.quantile(gap, 0.9) < 4.0 and gap.max() < 12.0
