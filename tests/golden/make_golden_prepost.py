"""Goldens for the feature-side pre/post steps.  Run in the BUILD container only
(needs /root/reference):   python tests/golden/make_golden_prepost.py

  norm_X   : the reference's own helpers/FeatureExtractor.Read_Landmarks_and_Normalizing_using_IPD
             (FeatureExtractor.py:30-66) applied to fake MediaPipe landmark objects, then
             torch.tensor(...).float() as FeatureExtractor.py:105
  deg3     : round(np.degrees(t.item()), 3) as NLML_HPE_Test.py:273 on float32 radians
  deg2_ema : round(..., 2) + the exponential smoothing block of generatePose_on_video.py:210-224 (alpha = 0.4),
             executed as the same Python statements
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from nlml_hpe_b200 import synthetic  # noqa: E402


def main():
    sys.path.insert(0, os.path.join(REF, "helpers"))
    import FeatureExtractor as FE
    art = np.load(f"{HERE}/shipped_artifacts.npz")
    rows = [art[f"optimized_{k}"][0:3] for k in ("yaw", "pitch", "roll")]
    n = 96
    feats = synthetic.make_features(n, art["W"], *rows, U_id=art["U_id"], seed=4321)
    rng = np.random.default_rng(99)
    # raw landmarks as MediaPipe would give them: image-normalised coordinates, float32 values
    scale = rng.uniform(0.05, 0.3, (n, 1, 1))
    shift = rng.uniform(0.3, 0.7, (n, 1, 3))
    raw = (feats.reshape(n, 468, 3).astype(np.float64) * scale + shift).astype(np.float32)
    raw[5] = 0.25                      # degenerate face: all landmarks equal -> ipd == 0 -> 1e-6 branch
    raw[6, 33] = raw[6, 263]           # eye corners coincide only
    norm = np.empty((n, 1404), np.float32)
    for i in range(n):
        lms = [types.SimpleNamespace(x=float(p[0]), y=float(p[1]), z=float(p[2])) for p in raw[i]]
        ref_point = lms[1]
        lst = FE.Read_Landmarks_and_Normalizing_using_IPD(lms, [ref_point.x, ref_point.y, ref_point.z], True)
        norm[i] = torch.tensor(lst[0:1404]).float().numpy()
    rad = rng.uniform(-1.2, 1.2, (257, 3)).astype(np.float32)
    rad[0] = [0.0, -0.0, 1e-7]
    rad[1] = [np.radians(12.3445), np.radians(-7.0005), np.radians(89.9995)]     # near rounding ties
    deg3 = np.array([[round(np.degrees(torch.tensor(v).item()), 3) for v in row] for row in rad], dtype=np.float64)
    alpha = 0.4
    sm = np.empty((len(rad), 3), np.float64)
    prediction_num = 0
    for t in range(len(rad)):
        prediction_num += 1
        predictions = [torch.tensor(v) for v in rad[t]]
        yaw, pitch, roll = round(np.degrees(predictions[0].item()), 2), round(np.degrees(predictions[1].item()), 2), round(np.degrees(predictions[2].item()), 2)
        if prediction_num <= 1:
            yaw_smoothed, pitch_smoothed, roll_smoothed = yaw, pitch, roll
        else:
            yaw_smoothed = alpha * yaw + (1 - alpha) * yaw_smoothed
            pitch_smoothed = alpha * pitch + (1 - alpha) * pitch_smoothed
            roll_smoothed = alpha * roll + (1 - alpha) * roll_smoothed
        sm[t] = [yaw_smoothed, pitch_smoothed, roll_smoothed]
    np.savez_compressed(f"{HERE}/prepost_golden.npz", raw=raw, norm_X=norm, rad=rad, deg3=deg3, deg2_ema=sm,
                        alpha=np.array(alpha), numpy_version=np.array(np.__version__))
    print("prepost_golden.npz written", raw.shape, norm.shape, "max|norm|", np.abs(norm[np.isfinite(norm).all(1)]).max())


if __name__ == "__main__":
    main()
