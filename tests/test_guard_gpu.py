"""Bounds checks of our own for the kernels that write through TMA or read strided inputs (compute-sanitizer is not
available on the GPU pool): outputs are placed inside NaN-filled buffers and must leave every guard element untouched;
inputs are views into NaN-filled buffers and must give the bits of the contiguous call (a stray read would poison the
result).  Row counts are ragged on purpose: not multiples of the 32-row store boxes or the 128-row tiles.
"""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _features(art, rows, n, seed=17):
    from nlml_hpe_b200 import synthetic
    return synthetic.make_features_torch(n, art["W"], *rows, U_id=art["U_id"], seed=seed, device="cuda")


@pytest.fixture(scope="module")
def model(state_dicts, cuda_lib):
    from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB
    return MB.build_combined_model(*state_dicts)


@pytest.fixture(scope="module")
def fitter(art, rows, cuda_lib):
    from nlml_hpe_b200.tucker import TuckerFitter
    return TuckerFitter(art["W"], *rows, device="cuda:0")


@pytest.mark.parametrize("n", [1, 31, 77, 129, 300, 4300])
def test_mlp_output_guards(model, art, rows, n):
    """nlml_mlp_forward_f32 writes rows [0, n) of its [n][3] output and nothing else; the internal activation planes
    leave through TMA boxes clipped at row n (the next call on a LARGER batch must not see leftovers either)."""
    from nlml_hpe_b200 import _lib
    X = _features(art, rows, n)
    ref = model.predict(X).clone()
    assert torch.isfinite(ref).all()
    buf = torch.full((n + 64, 3), float("nan"), device="cuda")
    plan = model._get_plan(0)
    _lib.check(plan.lib.nlml_mlp_forward_f32(plan.h, X.data_ptr(), n, X.shape[1], buf[32:].data_ptr(),
                                             torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(buf[32:32 + n], ref)
    assert torch.isnan(buf[:32]).all() and torch.isnan(buf[32 + n:]).all()


def test_mlp_input_guards(model, art, rows):
    """Rows and pad columns around the input view are NaN: the operand split reads exactly [n][1404]."""
    n = 333
    X = _features(art, rows, n)
    big = torch.full((n + 2, 1408), float("nan"), device="cuda")
    big[1:n + 1, :1404] = X
    assert torch.equal(model.predict(big[1:n + 1, :1404]), model.predict(X))
    big2 = torch.full((n + 2, 1411), float("nan"), device="cuda")     # unaligned rows: the scalar load path
    big2[1:n + 1, 3:1407] = X
    assert torch.equal(model.predict(big2[1:n + 1, 3:1407]), model.predict(X))


@pytest.mark.parametrize("kernel,n", [("tensor_core", 4300), ("tensor_core", 130), ("thread_per_sample", 300),
                                      ("warp_per_sample", 40), ("tensor_core_generic", 200)])
def test_tucker_output_and_input_guards(fitter, art, rows, kernel, n):
    """fit() into a strided window of a NaN-filled buffer, from a strided window of a NaN-filled buffer: same bits as
    the contiguous call, guards untouched.  4300 rows take the tensor-core projection (TMA loads with row pitch 1408,
    rows past n zero-filled by the tensor map) and the slab copy; 130 rows the in-kernel projection."""
    X = _features(art, rows, n, seed=5)
    ref = fitter.fit(X, 6, kernel=kernel).clone()
    assert torch.isfinite(ref).all()
    big = torch.full((n + 2, 1408), float("nan"), device="cuda")
    big[1:n + 1, :1404] = X
    out = torch.full((n + 32, 12), float("nan"), device="cuda")
    view = out[16:16 + n, :8]
    fitter.fit(big[1:n + 1, :1404], 6, kernel=kernel, out=view)
    torch.cuda.synchronize()
    assert torch.equal(view, ref)
    assert torch.isnan(out[:16]).all() and torch.isnan(out[16 + n:]).all() and torch.isnan(out[16:16 + n, 8:]).all()


def test_solve_guards(fitter, art, rows):
    n = 4300
    X = _features(art, rows, n, seed=6)
    ref = fitter.solve(X).clone()
    big = torch.full((n + 2, 1408), float("nan"), device="cuda")
    big[1:n + 1, :1404] = X
    out = torch.full((n + 32, 12), float("nan"), device="cuda")
    view = out[16:16 + n, :8]
    fitter.solve(big[1:n + 1, :1404], out=view)
    torch.cuda.synchronize()
    assert torch.equal(view, ref)
    assert torch.isnan(out[:16]).all() and torch.isnan(out[16 + n:]).all() and torch.isnan(out[16:16 + n, 8:]).all()
