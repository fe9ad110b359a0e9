"""The reference's scipy-Powell searches, restated in csrc/powell_math.h: TD_Tester.Test (TD_Tester.py:162-199) and
TD_Trainer.Train (TD_Trainer.py:232-351).  CPU tier: the host build of the very statements the CUDA kernels compile,
against outputs of the REAL reference (tests/golden/powell_golden.npz, trainer_golden.npz) -- bit for bit for Test(),
including the number of function evaluations.  GPU tier: the kernels through the C ABI and the Python entry points."""
import ctypes

import numpy as np
import pytest
import torch

DEG = 180.0 / np.pi
vp = ctypes.c_void_p


@pytest.fixture(scope="module")
def trainer_golden():
    import os
    from conftest import GOLDEN
    return dict(np.load(os.path.join(GOLDEN, "trainer_golden.npz")))


def _host_powell(lib, W, rows, X, exact):
    ranks, F = W.shape[:4], W.shape[4]
    W2 = np.ascontiguousarray(W.reshape(-1, F), dtype=np.float32)
    r = [np.ascontiguousarray(x, np.float64) for x in rows]
    X = np.ascontiguousarray(X, dtype=np.float32)
    n = len(X)
    P, fun, nfev = np.zeros((n, 3 + ranks[0])), np.zeros(n), np.zeros(n, np.int32)
    rc = lib.hostcheck_powell_tucker(vp(W2.ctypes.data), *[ctypes.c_int(int(v)) for v in ranks], ctypes.c_int(F), vp(r[0].ctypes.data),
                                     vp(r[1].ctypes.data), vp(r[2].ctypes.data), vp(X.ctypes.data), ctypes.c_int64(n), ctypes.c_int64(F),
                                     vp(P.ctypes.data), vp(fun.ctypes.data), vp(nfev.ctypes.data), ctypes.c_int(1 if exact else 0))
    assert rc == 0
    return P, fun, nfev


def test_powell_restatement_reproduces_the_reference_bit_for_bit(hostcheck, art, rows, X1k, powell_golden):
    """scipy's bracket / Brent / Powell and the reference's objective in its own operation order: same x, same f(x), same
    number of evaluations as TD_Tester.Test on the real reference (12 of the 96 goldens here; all 96 on the GPU)."""
    idx = powell_golden["idx"][[0, 3, 5, 17, 29, 41, 53, 60, 68, 77, 88, 95]]
    P, fun, nfev = _host_powell(hostcheck, art["W"], rows, X1k[idx], exact=True)
    assert np.array_equal(nfev, powell_golden["nfev"][idx])
    assert np.array_equal(P, powell_golden["p"][idx])
    assert np.array_equal(fun, powell_golden["loss"][idx])
    assert np.array_equal(np.degrees(P[:, :3]), powell_golden["deg"][idx])       # what Test() returns (:196-198)


def test_powell_restatement_on_the_folded_objective(hostcheck, art, rows, X1k, powell_golden):
    """The same search over the folded-Gram evaluation of the objective (float64; ~1e-14 of rounding noise instead of the
    reference's exact bits): the searches follow slightly different paths and end within Powell's own tolerance."""
    idx = powell_golden["idx"][:48]
    P, fun, nfev = _host_powell(hostcheck, art["W"], rows, X1k[idx], exact=False)
    gap = np.abs(np.degrees(P[:, :3]) - powell_golden["deg"][idx]).max(1)
    assert np.median(gap) < 2e-2 and np.abs(fun - powell_golden["loss"][idx]).max() < 1e-3


def _host_cosine(lib, U, w):
    U = np.ascontiguousarray(U, np.float64)
    w = np.ascontiguousarray(w, np.float64)
    nc = U.shape[1]
    init, out, fun, nfev = np.zeros((nc, 4)), np.zeros((nc, 4)), np.zeros(nc), np.zeros(nc, np.int32)
    assert lib.hostcheck_cosine_fit(vp(U.ctypes.data), U.shape[0], nc, vp(w.ctypes.data), vp(init.ctypes.data), vp(out.ctypes.data),
                                    vp(fun.ctypes.data), vp(nfev.ctypes.data)) == 0
    return out, init, fun


def _curve(P, w):
    w = np.radians(np.asarray(w, np.float64))
    return P[:, 0] * np.cos(P[:, 1] * w[:, None] + P[:, 2]) + P[:, 3]


@pytest.mark.parametrize("tag", ["shipped", "synth"])
def test_cosine_trainer_restatement_matches_the_reference(hostcheck, trainer_golden, art, tag):
    """TD_Trainer.Train: Fourier initial guess + Powell per column, against the real reference on the shipped factor
    matrices (whose result IS the shipped optimized_* rows) and on synthetic ones (21/17/13 bins, ranks 4/5/2)."""
    for k in ("yaw", "pitch", "roll"):
        U, w = trainer_golden[f"{tag}_{k}_U"], trainer_golden[f"{tag}_{k}_w"]
        out, init, fun = _host_cosine(hostcheck, U, w)
        # the shipped factor matrices are float32 and scipy.fft keeps single precision for them: the reference's initial
        # guess carries float32 rounding there; the float64 DFT here agrees to that level, and to 1e-9 on float64 input
        assert np.abs(init - trainer_golden[f"{tag}_{k}_init"]).max() < (5e-7 if tag == "shipped" else 1e-9), k
        ref = trainer_golden[f"{tag}_{k}_fit"]
        assert np.abs(out - ref).max() < 1e-6, (k, np.abs(out - ref).max())
        assert np.abs(_curve(out, w) - _curve(ref, w)).max() < (1e-6 if tag == "shipped" else 1e-7)
        if tag == "shipped":
            assert np.abs(out - art[f"optimized_{k}"]).max() < 1e-6       # the rows the hot path loads (TD_Inference.py:43-45)


# ---- GPU tier ----------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_powell_kernel_reproduces_the_reference_on_all_96_goldens(art, rows, X1k, powell_golden, cuda_lib):
    from nlml_hpe_b200.tucker import TuckerFitter
    fit = TuckerFitter(art["W"], *rows, device="cuda:0")
    idx = powell_golden["idx"]
    P, fun, nfev = fit.powell(torch.from_numpy(X1k[idx]).cuda(), return_info=True)
    P, fun, nfev = P.cpu().numpy(), fun.cpu().numpy(), nfev.cpu().numpy()
    same = (nfev == powell_golden["nfev"])
    gap = np.abs(np.degrees(P[:, :3]) - powell_golden["deg"]).max(1)
    print(f"Powell kernel vs reference: {same.mean() * 100:.1f} % identical evaluation counts, {np.mean(gap == 0) * 100:.1f} % "
          f"bit-identical angles, max gap {gap.max():.3e} deg")
    # measured on B200: 100 % / 100 % / 0.0 (the device's float64 cos may differ from numpy's in the last bit, but the
    # objective rounds the factors to float32, which hides that except exactly on a rounding boundary)
    assert same.all() and np.array_equal(P, powell_golden["p"]) and np.array_equal(fun, powell_golden["loss"])


@pytest.mark.gpu
def test_Test_entry_point_returns_the_reference_result(art, rows, X1k, powell_golden, cuda_lib):
    """TD_Tester.Test with its default solver: the tuple the reference returns (degrees as numpy float64, the given u_id)."""
    from nlml_hpe_b200 import TD_Tester
    assert TD_Tester.TEST_SOLVER == "powell"
    for i in (0, 1, 2, 11):
        y, p, r, u = TD_Tester.Test(art["W"], torch.from_numpy(X1k[i]), 5, *rows, None, None, None, None)
        assert u is None
        assert max(abs(y - powell_golden["deg"][i, 0]), abs(p - powell_golden["deg"][i, 1]), abs(r - powell_golden["deg"][i, 2])) < 1e-9


@pytest.mark.gpu
def test_powell_kernel_other_ranks_and_edges(cuda_lib, hostcheck):
    """Run-time ranks and a feature count that is not a multiple of anything: kernel vs the host build of the same text."""
    from nlml_hpe_b200 import synthetic
    from nlml_hpe_b200.tucker import TuckerFitter
    ranks, F = (4, 3, 2, 4), 203
    G = synthetic.synthetic_core(ranks, F, seed=3, std=1.0)
    rws = [synthetic.synthetic_cos_params(r, 50 + i) for i, r in enumerate(ranks[1:])]
    X = synthetic.make_features(9, G, *rws, U_id=None, seed=6)
    X[4] = 0.0
    fit = TuckerFitter(G, *rws, device="cuda:0")
    P, fun, nfev = fit.powell(torch.from_numpy(X).cuda(), return_info=True)
    Ph, funh, nfevh = _host_powell(hostcheck, G, rws, X, exact=True)
    assert (nfev.cpu().numpy() == nfevh).mean() >= 0.8
    assert np.abs(P.cpu().numpy() - Ph).max() < 1e-6 or np.median(np.abs(P.cpu().numpy() - Ph)) == 0.0
    assert fit.powell(torch.from_numpy(X[:0]).cuda()).shape == (0, 7)
    with pytest.raises(Exception):
        fit.powell(X)                                   # host array into the device entry point


@pytest.mark.gpu
def test_TD_Trainer_Train_and_compute_W(trainer_golden, art, capsys, cuda_lib):
    from nlml_hpe_b200 import TD_Trainer
    for tag in ("shipped", "synth"):
        sets = [(trainer_golden[f"{tag}_{k}_U"], trainer_golden[f"{tag}_{k}_w"]) for k in ("yaw", "pitch", "roll")]
        oy, op_, or_ = TD_Trainer.Train(*sets)
        assert "Optimal parameters for yaw:" in capsys.readouterr().out
        init = TD_Trainer.estimate_init_Fourier_Trans(*sets)
        for k, got, i0, (U, w) in zip(("yaw", "pitch", "roll"), (oy, op_, or_), init, sets):
            ref = trainer_golden[f"{tag}_{k}_fit"]
            assert np.abs(i0 - trainer_golden[f"{tag}_{k}_init"]).max() < (5e-7 if tag == "shipped" else 1e-9)
            assert np.abs(_curve(got, w) - _curve(ref, w)).max() < 1e-6, (tag, k)
            assert np.abs(got - ref).max() < 1e-5, (tag, k, np.abs(got - ref).max())
    # W = core x_5 U_feat (TD_main.py:231-238)
    rng = np.random.default_rng(0)
    core = rng.standard_normal((5, 3, 3, 3, 40)).astype(np.float32)
    Uf = rng.standard_normal((203, 40)).astype(np.float32)
    W = TD_Trainer.compute_W(core, Uf)
    ref = np.tensordot(core.astype(np.float64), Uf.T.astype(np.float64), axes=(4, 0))
    assert W.shape == (5, 3, 3, 3, 203) and np.abs(W - ref).max() < 1e-4
