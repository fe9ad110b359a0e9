"""Deterministic synthetic inputs for the NLML_HPE inference hot path.

The reference ships no throughput benchmark and its feature extractor
(MediaPipe FaceMesh, helpers/FeatureExtractor.py:71-111) cannot run here, so
both halves of the path are driven with synthetic landmark-feature batches
``X[N, 1404]`` that lie on the Tucker manifold plus a little noise
(SURVEY.md section 8d).  Everything here is numpy ``default_rng`` (PCG64), so
the CPU container and the GPU box regenerate bit-identical inputs from a seed
and only the *outputs* need to be committed as golden fixtures.

Nothing in this module is on the timed path.
"""
from __future__ import annotations

import numpy as np

# angle ranges of the training grid, configs/config_TD_main.yaml:15-28 (degrees)
ANGLE_RANGE_DEG = {"yaw": 50.0, "pitch": 40.0, "roll": 30.0}

# hidden widths of the reference networks (NLML_HPE_Model_Builder.py:33-53, 76-92)
ENCODER_WIDTHS = (1024, 512, 256, 128, 64)
HEAD_WIDTHS = (128, 256, 128, 64)


def cos_rows(w, params):
    """Factor-row model a*cos(b*w+c)+d of TD_Tester.py:25-28 for an array of angles.

    w: [...] radians; params: [R,4] rows (a,b,c,d).  Returns [..., R] float64.
    """
    w = np.asarray(w, dtype=np.float64)[..., None]
    p = np.asarray(params, dtype=np.float64)
    return p[:, 0] * np.cos(p[:, 1] * w + p[:, 2]) + p[:, 3]


def make_features(n, W, params_y, params_p, params_r, U_id=None, seed=1234, noise=0.01,
                  return_truth=False):
    """On-manifold feature vectors x = W x1 u x2 c(yaw) x3 c(pitch) x4 c(roll) + N(0, noise^2).

    W: f32 [R_id,R_y,R_p,R_r,F].  U_id: optional [K,R_id] pool of identity rows
    (one is drawn per sample); when None the identity is drawn from the measured
    column statistics of the shipped U_id.
    Returns X f32 [n,F] (and the true angles in degrees + identity rows).
    """
    rng = np.random.default_rng(seed)
    W = np.asarray(W, dtype=np.float32)
    r_id, r_y, r_p, r_r, F = W.shape
    yaw = rng.uniform(-ANGLE_RANGE_DEG["yaw"], ANGLE_RANGE_DEG["yaw"], n)
    pitch = rng.uniform(-ANGLE_RANGE_DEG["pitch"], ANGLE_RANGE_DEG["pitch"], n)
    roll = rng.uniform(-ANGLE_RANGE_DEG["roll"], ANGLE_RANGE_DEG["roll"], n)
    if U_id is not None:
        U_id = np.asarray(U_id, dtype=np.float64)
        u = U_id[rng.integers(0, U_id.shape[0], n)][:, :r_id]
    else:
        u = rng.normal(0.0, 0.0248, (n, r_id))
        u[:, 0] = rng.normal(0.0248, 0.0007, n)
    cy = cos_rows(np.radians(yaw), params_y)
    cp = cos_rows(np.radians(pitch), params_p)
    cr = cos_rows(np.radians(roll), params_r)
    z = np.einsum("ni,nj,nk,nl->nijkl", u, cy, cp, cr).reshape(n, -1)
    X = z @ W.reshape(-1, F).astype(np.float64)
    X += rng.normal(0.0, noise, (n, F))
    X = X.astype(np.float32)
    if return_truth:
        return X, np.stack([yaw, pitch, roll], 1), u.astype(np.float32)
    return X


def synthetic_core(ranks=(5, 3, 3, 3), F=1404, seed=7, std=6.5):
    """Synthetic Tucker core 'of the configured rank' (BASELINE.json configs 2/4/5).

    Shape from configs/config_TD_main.yaml:8-13; entries N(0, std^2) with std matched
    to the shipped W.
    """
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((*ranks, F)) * std).astype(np.float32)


def synthetic_cos_params(rank, seed, base=None):
    """[rank,4] rows (a,b,c,d); the first rows are `base` (shipped fit) when given."""
    rng = np.random.default_rng(seed)
    out = np.empty((rank, 4), dtype=np.float64)
    out[:, 0] = rng.uniform(-1.0, 1.0, rank)
    out[:, 1] = rng.uniform(0.5, 3.0, rank)
    out[:, 2] = rng.uniform(-np.pi, np.pi, rank)
    out[:, 3] = rng.uniform(-0.2, 0.2, rank)
    if base is not None:
        k = min(rank, len(base))
        out[:k] = np.asarray(base, dtype=np.float64)[:k]
    return out


def _round_sig(a, digits=3):
    """Round to `digits` significant digits (makes the calibration immune to last-bit BLAS differences)."""
    a = np.asarray(a, dtype=np.float64)
    mag = np.where(a == 0, 1.0, 10.0 ** np.floor(np.log10(np.abs(np.where(a == 0, 1.0, a)))))
    return np.round(a / mag, digits - 1) * mag


def synthetic_encoder_state_dict(W, params_y, params_p, params_r, U_id=None, seed=0, input_size=1404):
    """Stand-in for the missing models/Encoder.pth (.MISSING_LARGE_BLOBS:4).

    Same keys/shapes as LandmarkEncoder(1404, [(1,3)]*3).state_dict()
    (NLML_HPE_Model_Builder.py:33-53): encoder.{0,2,4,6,8,10}.{weight,bias}.
    Hidden layers start from nn.Linear's default U(-1/sqrt(in), 1/sqrt(in)) range and are
    then calibrated on a fixed synthetic batch so every pre-activation has zero mean and unit
    spread (otherwise a random ReLU stack maps all poses to nearly the same latent and the
    parity test would only exercise one point of the heads).  The last layer is centred and
    scaled so the 9 latents sweep the range of the cosine features the shipped heads were
    trained on.  Returns numpy f32 arrays.
    """
    rng = np.random.default_rng(seed)
    widths = (input_size,) + ENCODER_WIDTHS
    h = make_features(256, W, np.asarray(params_y)[:W.shape[1]], np.asarray(params_p)[:W.shape[2]],
                      np.asarray(params_r)[:W.shape[3]], U_id=U_id, seed=777).astype(np.float64)
    sd = {}
    for li in range(len(ENCODER_WIDTHS)):
        fan_in, fan_out = widths[li], widths[li + 1]
        bound = 1.0 / np.sqrt(fan_in)
        w = rng.uniform(-bound, bound, (fan_out, fan_in))
        b = rng.uniform(-bound, bound, fan_out)
        pre = h @ w.T + b
        mean, std = _round_sig(pre.mean(0)), _round_sig(pre.std(0) + 1e-12)
        w, b = w / std[:, None], (b - mean) / std
        w, b = w.astype(np.float32), b.astype(np.float32)
        sd[f"encoder.{2 * li}.weight"], sd[f"encoder.{2 * li}.bias"] = w, b
        pre = h @ w.astype(np.float64).T + b
        h = np.tanh(pre) if li == len(ENCODER_WIDTHS) - 1 else np.maximum(pre, 0.0)
    centres, halves = [], []
    for name, P in (("yaw", params_y), ("pitch", params_p), ("roll", params_r)):
        w = np.radians(np.linspace(-ANGLE_RANGE_DEG[name], ANGLE_RANGE_DEG[name], 201))
        c = cos_rows(w, np.asarray(P)[:3])
        centres.append(0.5 * (c.max(0) + c.min(0)))
        halves.append(0.5 * (c.max(0) - c.min(0)))
    centres = np.concatenate(centres)
    halves = np.concatenate(halves)
    w5 = rng.uniform(-1.0, 1.0, (9, ENCODER_WIDTHS[-1]))
    spread = _round_sig((h @ w5.T).std(0) + 1e-12)
    w5 *= (halves / (2.0 * spread))[:, None]
    sd["encoder.10.weight"] = w5.astype(np.float32)
    sd["encoder.10.bias"] = centres.astype(np.float32)
    return sd


def make_features_torch(n, W, params_y, params_p, params_r, U_id=None, seed=1234, noise=0.01, device="cuda",
                        out=None, chunk=65536):
    """Device-side twin of make_features for the 1M-sample workloads (same distribution, torch RNG).

    Generated chunk by chunk straight into `out` ([n,F] float32 on `device`), so the 5.6 GB batch of
    BASELINE.json configs 3/4 never exists on the host.
    """
    import torch

    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    Wt = torch.as_tensor(np.asarray(W, dtype=np.float32), device=dev)
    r_id, r_y, r_p, r_r, F = Wt.shape
    W2 = Wt.reshape(-1, F)
    rows = [torch.as_tensor(np.asarray(p, dtype=np.float32)[:r], device=dev)
            for p, r in ((params_y, r_y), (params_p, r_p), (params_r, r_r))]
    Uid = None if U_id is None else torch.as_tensor(np.asarray(U_id, dtype=np.float32)[:, :r_id], device=dev)
    if out is None:
        out = torch.empty((n, F), dtype=torch.float32, device=dev)
    lims = [ANGLE_RANGE_DEG[k] * np.pi / 180.0 for k in ("yaw", "pitch", "roll")]
    for s0 in range(0, n, chunk):
        m = min(chunk, n - s0)
        feats = []
        for lim, r in zip(lims, rows):
            w = (torch.rand(m, generator=gen, device=dev) * 2 - 1) * lim
            feats.append(r[:, 0] * torch.cos(r[:, 1] * w[:, None] + r[:, 2]) + r[:, 3])
        if Uid is not None:
            u = Uid[torch.randint(0, Uid.shape[0], (m,), generator=gen, device=dev)]
        else:
            u = torch.randn(m, r_id, generator=gen, device=dev) * 0.0248
            u[:, 0] = 0.0248 + torch.randn(m, generator=gen, device=dev) * 0.0007
        z = torch.einsum("ni,nj,nk,nl->nijkl", u, *feats).reshape(m, -1)
        blk = out[s0:s0 + m]
        torch.matmul(z, W2, out=blk)
        blk.add_(torch.randn(m, F, generator=gen, device=dev), alpha=noise)
    return out
