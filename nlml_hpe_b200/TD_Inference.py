"""Drop-in for the reference's TD_Inference.py (single-image Tucker-fit CLI), B200-backed.

Reference flow (/root/reference/TD_Inference.py:20-76): --image_path -> MediaPipe FaceMesh ->
FE.get_feature_vector (f32[1404]) -> np.load Trained_data.npz / Factor_Matrices.npz -> TD_Tester.Test
-> print three lines.  Feature extraction is outside the hot path; when
MediaPipe (and the reference's helpers package) are importable they are used, otherwise pass a
pre-extracted feature vector with --features_npy.
"""
from __future__ import annotations

import argparse
import warnings

import numpy as np
import torch

from . import TD_Tester
from .config import load_tucker_artifacts


def _extract(image_path):
    try:
        import mediapipe as mp
        from helpers import FeatureExtractor as FE  # the reference's own helper, on the user's PYTHONPATH
    except ImportError as e:
        raise RuntimeError("feature extraction needs mediapipe and the reference's helpers/ package "
                           f"({e}); pass --features_npy with a saved float32[1404] vector instead") from e
    face_mesh = mp.solutions.face_mesh.FaceMesh(static_image_mode=True, max_num_faces=1,
                                                min_detection_confidence=0.5, min_tracking_confidence=0.5)
    return FE.get_feature_vector(face_mesh, image_path, normalize=True)


def inference(argv=None):
    warnings.filterwarnings("ignore", category=RuntimeWarning)
    parser = argparse.ArgumentParser(description="Inference the head pose of single input image")
    group = parser.add_mutually_exclusive_group(required=True)
    group.add_argument("--image_path", type=str, help="Path to input image")
    group.add_argument("--features_npy", type=str, help="Path to a saved float32[1404] landmark feature vector")
    args = parser.parse_args(argv)

    x = _extract(args.image_path) if args.image_path else torch.from_numpy(np.load(args.features_npy).astype(np.float32))
    art = load_tucker_artifacts()
    est_w_y, est_w_p, est_w_r, _ = TD_Tester.Test(art["W"], x, art["u_id_shape"], art["optimized_yaw"][0:3, :],
                                                  art["optimized_pitch"][0:3, :], art["optimized_roll"][0:3, :],
                                                  None, None, None, None)
    print(f"Estimated yaw in degree = {est_w_y:.2f}")
    print(f"Estimated pitch in degree = {est_w_p:.2f}")
    print(f"Estimated roll in degree = {est_w_r:.2f}")
    return est_w_y, est_w_p, est_w_r


if __name__ == "__main__":
    inference()
