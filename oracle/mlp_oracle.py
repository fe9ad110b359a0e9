"""CPU oracle for the Encoder + yaw/pitch/roll MLP-heads half of the hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this; the product (nlml_hpe_b200/) never does.

Parity status: PINNED -- `forward` is checked in tests/test_oracle.py against the real
reference classes (LandmarkEncoder / AnglePredictionNetwork / CombinedAnglePredictionModel,
eager and torch.jit.script) run in the build container by tests/golden/make_golden.py on the
shipped models/{yaw,pitch,roll}_network.pth plus the synthetic encoder (models/Encoder.pth is
missing from the checkout, .MISSING_LARGE_BLOBS:4).  Outputs committed under tests/golden/.

What is restated (file:line into /root/reference/NLML_HPE_Model_Builder.py):
  encoder stack   :33-53   Linear 1404-1024-512-256-128-64-9, ReLU x4, Tanh, none
  latent split    :55-68   latent[:,0:3], [:,3:6], [:,6:9]  (matrix_dims [(1,3)]*3, squeeze(1) :118-120)
  head stack      :76-92   Linear 3-128-256-128-64-1, ReLU x4, none
  combined        :115-126 returns (yaw, pitch, roll), each [B,1], radians
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

ENCODER_KEYS = [f"encoder.{i}" for i in (0, 2, 4, 6, 8, 10)]
HEAD_KEYS = [f"model.{i}" for i in (0, 2, 4, 6, 8)]


def _t(a, dtype):
    return torch.as_tensor(np.asarray(a)).to(dtype)


def forward(encoder_sd, yaw_sd, pitch_sd, roll_sd, X, dtype=torch.float32, threads=None):
    """X [B,1404] -> angles [B,3] (yaw, pitch, roll) radians, on CPU in `dtype`.

    dtype=float32 is the reference arithmetic (torch.nn.Linear on CPU); float64 gives the
    error-budget reference used to state tolerances.
    """
    # As a CHECKER (threads=None) the forward runs single-threaded: on the GPU boxes' host the multi-threaded f32 GEMM
    # of torch CPU was observed to return slightly different rows (2.9e-3 degrees on every other row) in about one
    # process out of ten (scripts/dbg_mlp_chunk2.py), which made a parity test flaky.  bench.py's CPU baseline passes
    # `threads` explicitly and keeps all host threads.
    prev_threads = torch.get_num_threads()
    torch.set_num_threads(1 if threads is None else threads)
    try:
        return _forward_impl(encoder_sd, yaw_sd, pitch_sd, roll_sd, X, dtype)
    finally:
        torch.set_num_threads(prev_threads)


def _forward_impl(encoder_sd, yaw_sd, pitch_sd, roll_sd, X, dtype):
    with torch.no_grad():
        h = _t(X, dtype)
        for li, key in enumerate(ENCODER_KEYS):
            h = F.linear(h, _t(encoder_sd[key + ".weight"], dtype), _t(encoder_sd[key + ".bias"], dtype))
            if li < 4:
                h = torch.relu(h)
            elif li == 4:
                h = torch.tanh(h)
        outs = []
        for hi, sd in enumerate((yaw_sd, pitch_sd, roll_sd)):
            width = _t(sd["model.0.weight"], dtype).shape[1]
            g = h[:, hi * width:(hi + 1) * width]
            for li, key in enumerate(HEAD_KEYS):
                g = F.linear(g, _t(sd[key + ".weight"], dtype), _t(sd[key + ".bias"], dtype))
                if li < 4:
                    g = torch.relu(g)
            outs.append(g)
        return torch.cat(outs, 1).numpy()


def latent(encoder_sd, X, dtype=torch.float32):
    """Encoder output [B,9] only (for intermediate-stage checks); single-threaded like the checker forward."""
    prev_threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        with torch.no_grad():
            h = _t(X, dtype)
            for li, key in enumerate(ENCODER_KEYS):
                h = F.linear(h, _t(encoder_sd[key + ".weight"], dtype), _t(encoder_sd[key + ".bias"], dtype))
                if li < 4:
                    h = torch.relu(h)
                elif li == 4:
                    h = torch.tanh(h)
            return h.numpy()
    finally:
        torch.set_num_threads(prev_threads)


# ---- feature-side pre / post steps of the reference's callers ----------------------
# PINNED: tests/golden/prepost_golden.npz holds outputs of the reference's own
# helpers/FeatureExtractor.Read_Landmarks_and_Normalizing_using_IPD run in the build container
# (tests/golden/make_golden_prepost.py) and of the callers' post-processing statements executed verbatim as Python.

LEFT_EYE, RIGHT_EYE, NOSE = 33, 263, 1    # helpers/FeatureExtractor.py:35-36, :89


def ipd_normalize(raw):
    """helpers/FeatureExtractor.py:30-66 + :89-90 + :105 for a batch: raw [N,468,3] (or [N,1404]) MediaPipe
    landmarks -> float32 features [N,1404]: subtract the nose tip (landmark 1), divide by the inter-pupillary
    distance ||lm33 - lm263|| (1e-6 when zero), all in float64 as the Python floats of the reference, then
    `.float()`."""
    raw = np.asarray(raw, dtype=np.float64).reshape(len(raw), -1, 3)
    out = np.empty(raw.shape, dtype=np.float64)
    for n in range(raw.shape[0]):
        d = raw[n, LEFT_EYE] - raw[n, RIGHT_EYE]
        ipd = np.linalg.norm(d)
        if ipd == 0:
            ipd = 1e-6
        out[n] = (raw[n] - raw[n, NOSE]) / ipd
    return out.reshape(len(raw), -1).astype(np.float32)


def degrees_round(rad, decimals=3):
    """round(np.degrees(t.item()), decimals) per angle (NLML_HPE_Test.py:273 decimals=3,
    generatePose_on_video.py:210 decimals=2): float32 radians -> float64 degrees."""
    rad = np.asarray(rad, dtype=np.float32)
    return np.array([[round(np.degrees(float(v)), decimals) for v in row] for row in rad], dtype=np.float64)


def ema(deg, alpha=0.4):
    """generatePose_on_video.py:215-224: s_0 = y_0; s_t = alpha*y_t + (1-alpha)*s_{t-1}, Python floats."""
    deg = np.asarray(deg, dtype=np.float64)
    out = np.empty_like(deg)
    for j in range(deg.shape[1]):
        s = 0.0
        for t in range(deg.shape[0]):
            y = float(deg[t, j])
            s = y if t == 0 else alpha * y + (1 - alpha) * s
            out[t, j] = s
    return out
