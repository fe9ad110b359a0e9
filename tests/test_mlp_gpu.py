"""Parity of the CUDA Encoder+heads forward (through the C ABI) with the reference goldens and the oracle.

Tolerance (BASELINE.json north_star): <= 1e-3 degrees on the predicted Euler angles.
"""
import numpy as np
import pytest
import torch

from oracle import mlp_oracle

pytestmark = pytest.mark.gpu
DEG = 180.0 / np.pi
TOL_DEG = 1e-3


@pytest.fixture(scope="module", params=["tensor_core", "fp32"])
def model(request, state_dicts, cuda_lib):
    """Both kernel generations must meet the same parity bar: the tcgen05 chain (default) and the FP32
    CUDA-core chain it is validated against."""
    from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB
    m = MB.build_combined_model(*state_dicts)
    m.set_path(request.param)
    return m


def _gpu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_golden_1k(model, X1k, mlp_golden):
    """BASELINE.json config 1: shipped heads (+ synthetic encoder) on 1k synthetic feature vectors."""
    with torch.no_grad():
        yaw, pitch, roll = model(_gpu(X1k))
    assert yaw.shape == pitch.shape == roll.shape == (1000, 1) and yaw.is_cuda
    out = torch.cat([yaw, pitch, roll], 1).cpu().numpy()
    assert np.abs(out - mlp_golden["angles_jit"]).max() * DEG < TOL_DEG
    assert np.abs(out - mlp_golden["angles_f64"]).max() * DEG < TOL_DEG   # and no further from exact than the reference


def test_latent_stage(model, X1k, mlp_golden):
    lat = model.latent(_gpu(X1k)).cpu().numpy()
    assert np.abs(lat - mlp_golden["latent"]).max() < 2e-5


def test_edge_inputs(model, mlp_golden):
    out = model.predict(_gpu(mlp_golden["edge_X"])).cpu().numpy()
    assert np.abs(out - mlp_golden["edge_angles"]).max() * DEG < TOL_DEG


def test_reference_style_batch1_loop(model, X1k, mlp_golden):
    """NLML_HPE_Test.py:326-328: one sample per call, .item(), degrees rounded to 3 decimals."""
    from nlml_hpe_b200.NLML_HPE_Test import predict_degrees
    for i in range(8):
        x = torch.from_numpy(X1k[i]).unsqueeze(0).float().to("cuda")
        with torch.no_grad():
            y, p, r = model(x)
        got = np.array([y.item(), p.item(), r.item()])
        assert np.abs(got - mlp_golden["angles_b1"][i]).max() * DEG < TOL_DEG
    tuples, keep = predict_degrees(model, np.concatenate([X1k[:4], np.zeros((1, 1404), np.float32)]))
    assert len(tuples) == 4 and keep.tolist() == [True] * 4 + [False]
    want = np.round(np.degrees(mlp_golden["angles_jit"][:4].astype(np.float64)), 3)
    assert np.abs(np.array(tuples) - want).max() <= 2e-3


@pytest.mark.parametrize("n", [0, 1, 2, 127, 128, 129, 1000])
def test_ragged_batch_sizes(model, X1k, n):
    full = model.predict(_gpu(X1k))
    part = model.predict(_gpu(X1k[:n]))
    assert part.shape == (n, 3)
    if n:
        assert (part - full[:n]).abs().max().item() * DEG < 1e-4


def test_chunk_boundaries_and_oracle_fresh_seed(model, state_dicts, art, rows):
    """More samples than one internal chunk (16384), fresh seed, checked against the oracle on a subset."""
    from nlml_hpe_b200 import synthetic
    X = synthetic.make_features(2048, art["W"], *rows, U_id=art["U_id"], seed=99)
    big = _gpu(X).repeat(9, 1)[:16384 + 777].contiguous()
    out = model.predict(big).cpu().numpy()
    ref = mlp_oracle.forward(*state_dicts, X)
    assert np.abs(out[:2048] - ref).max() * DEG < TOL_DEG
    assert np.abs(out[16384:16384 + 777] - ref[:777]).max() * DEG < TOL_DEG


def test_host_path_equals_device_path(model, X1k):
    dev = model.predict(_gpu(X1k)).cpu().numpy()
    host = model.predict_host(X1k)
    assert np.array_equal(dev, host)
    y, p, r = model(torch.from_numpy(X1k))          # CPU tensor in -> CPU tensors out, through the GPU
    assert not y.is_cuda and np.array_equal(torch.cat([y, p, r], 1).numpy(), dev)


def test_tensor_core_chain_matches_fp32_chain(state_dicts, X1k, cuda_lib):
    from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB
    m = MB.build_combined_model(*state_dicts)
    x = _gpu(X1k)
    m.set_path("fp32")
    a, la = m.predict(x), m.latent(x)
    m.set_path("tensor_core")
    b, lb = m.predict(x), m.latent(x)
    assert (la - lb).abs().max().item() < 2e-5
    assert (a - b).abs().max().item() * DEG < TOL_DEG


def test_random_weights_and_inputs(cuda_lib):
    """Not just the shipped weights: random nets, N(0,1) inputs, checked against the oracle."""
    from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB
    torch.manual_seed(3)
    enc = MB.LandmarkEncoder(1404, [(1, 3)] * 3)
    heads = [MB.AnglePredictionNetwork(3) for _ in range(3)]
    model = MB.CombinedAnglePredictionModel(enc, *heads).eval()
    sds = [{k: v.numpy() for k, v in m.state_dict().items()} for m in (enc, *heads)]
    X = torch.randn(513, 1404)
    out = model.predict(X.cuda()).cpu().numpy()
    ref = mlp_oracle.forward(*sds, X.numpy())
    assert np.abs(out - ref).max() < 1.7e-5   # 1e-3 degrees (BASELINE.json north_star) in radians


def test_full_size_properties_1M(model, art, rows):
    """BASELINE.json config 3 size (1M vectors): periodic input => periodic output."""
    from nlml_hpe_b200 import synthetic
    n = 1_000_000
    base = _gpu(synthetic.make_features(4096, art["W"], *rows, U_id=art["U_id"], seed=55))
    X = base.repeat(n // 4096 + 1, 1)[:n].contiguous()
    out = model.predict(X)
    torch.cuda.synchronize()
    small = model.predict(base)
    assert out.shape == (n, 3) and torch.isfinite(out).all()
    assert (out[:4096] - small).abs().max().item() * DEG < 1e-4
    assert (out[4096 * 200:4096 * 201] - small).abs().max().item() * DEG < 1e-4


# ---- feature-side pre / post steps fused around the forward ----

@pytest.fixture(scope="module")
def tc_model(state_dicts, cuda_lib):
    from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB
    return MB.build_combined_model(*state_dicts)


def _same_bits(a, b):
    """Bit-for-bit equality, NaN included (the two degenerate faces of the fixture, IPD = 0, are outside FP16's range and
    come back NaN on the tensor-core path: see test_out_of_range_inputs_fail_loudly)."""
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    return a.shape == b.shape and np.array_equal(np.ascontiguousarray(a).view(np.int32), np.ascontiguousarray(b).view(np.int32))


def test_fused_ipd_normalisation_is_bit_exact(tc_model, prepost_golden):
    """Raw landmarks through the fused load stage == the reference-normalised features through the plain forward,
    bit for bit (so the float64 normalisation on the device reproduces FeatureExtractor.py:30-66 + .float())."""
    g = prepost_golden
    raw = _gpu(g["raw"])                                   # [96,468,3]
    fused = tc_model.predict_landmarks(raw).cpu().numpy()
    plain = tc_model.predict(_gpu(g["norm_X"])).cpu().numpy()
    assert _same_bits(fused, plain)
    assert _same_bits(tc_model.predict_landmarks(raw.reshape(96, 1404)).cpu().numpy(), plain)
    assert _same_bits(tc_model.predict_landmarks(raw[:1]).cpu().numpy(), plain[:1])
    assert _same_bits(tc_model.predict_landmarks_host(g["raw"]), plain)                     # host buffers
    # and against the oracle end to end (normalise on the CPU in float64, forward in float32)
    ok = np.abs(g["norm_X"]).max(1) < 10.0                # the two degenerate faces are outside the 1e-3 degree budget's range
    ref = mlp_oracle.forward(*[s for s in tc_model_state(tc_model)], mlp_oracle.ipd_normalize(g["raw"])[ok])
    assert np.abs(fused[ok] - ref).max() * DEG < TOL_DEG


def tc_model_state(m):
    enc = {k: v.detach().cpu().numpy() for k, v in m.encoder.state_dict().items()}
    heads = [{k: v.detach().cpu().numpy() for k, v in h.state_dict().items()} for h in (m.yaw_network, m.pitch_network, m.roll_network)]
    return enc, heads[0], heads[1], heads[2]


def test_fused_ipd_normalisation_large_batch_and_strides(tc_model, prepost_golden):
    g = prepost_golden
    base_raw, base_norm = _gpu(g["raw"].reshape(96, 1404)), _gpu(g["norm_X"])
    reps = 1700                                            # 163 200 rows: crosses two chunk boundaries
    fused = tc_model.predict_landmarks(base_raw.repeat(reps, 1))
    plain = tc_model.predict(base_norm.repeat(reps, 1))
    assert _same_bits(fused, plain)
    wide = torch.zeros((96, 1500), device="cuda")
    wide[:, 3:1407] = base_raw                             # unaligned rows, row stride != 1404
    assert _same_bits(tc_model.predict_landmarks(wide[:, 3:1407]), plain[:96])


def test_fused_ipd_normalisation_fp32_path_fails_loudly(state_dicts, prepost_golden, cuda_lib):
    from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB
    m = MB.build_combined_model(*state_dicts)
    m.set_path("fp32")
    with pytest.raises(Exception, match="tensor-core"):
        m.predict_landmarks(_gpu(prepost_golden["raw"]))


def test_degrees_rounding_and_ema_are_bit_exact(tc_model, prepost_golden):
    g = prepost_golden
    rad = _gpu(g["rad"])
    assert np.array_equal(tc_model.to_degrees(rad, 3).cpu().numpy(), g["deg3"])                      # NLML_HPE_Test.py:273
    assert np.array_equal(tc_model.to_degrees(rad, 2, ema_alpha=float(g["alpha"])).cpu().numpy(), g["deg2_ema"])   # video demo
    assert np.array_equal(tc_model.to_degrees(rad, 2).cpu().numpy(), mlp_oracle.degrees_round(g["rad"], 2))
    assert tc_model.to_degrees(rad[:0], 3).shape == (0, 3)
    big = torch.from_numpy(np.random.default_rng(3).uniform(-1.5, 1.5, (200_000, 3)).astype(np.float32)).cuda()
    out = tc_model.to_degrees(big, 3).cpu().numpy()
    assert np.array_equal(out[:5000], mlp_oracle.degrees_round(big[:5000].cpu().numpy(), 3))
    assert np.array_equal(out, np.round(np.degrees(big.cpu().numpy().astype(np.float64)), 3))
    with pytest.raises(Exception):
        tc_model.to_degrees(rad, 3, ema_alpha=1.5)


def test_out_of_range_inputs_fail_loudly(tc_model, state_dicts, X1k):
    """Inputs beyond FP16's range (e.g. raw pixel coordinates times a large factor): the tensor-core path returns
    NON-FINITE rows for them -- not finite-but-wrong angles (the ReLUs propagate NaN, as torch.relu does) -- and leaves
    the in-range rows of the same batch intact; the FP32 path stays finite like the reference's FP32 forward."""
    X = X1k[:256].copy()
    X[::2] *= 3.0e5                                        # every other row far outside +-65504
    out = tc_model.predict(_gpu(X)).cpu().numpy()
    assert (~np.isfinite(out[::2])).any(axis=1).all()      # every out-of-range row is flagged by a non-finite value
    ref_in = mlp_oracle.forward(*state_dicts, X[1::2])
    assert np.abs(out[1::2] - ref_in).max() * DEG < TOL_DEG
    tc_model.set_path("fp32")
    try:
        full = tc_model.predict(_gpu(X)).cpu().numpy()
    finally:
        tc_model.set_path("tensor_core")
    ref = mlp_oracle.forward(*state_dicts, X)
    assert np.isfinite(ref).all() and np.isfinite(full).all()
    assert np.abs(full[1::2] - ref[1::2]).max() * DEG < TOL_DEG
