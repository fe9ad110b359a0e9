"""Drop-in for the forward-call part of the reference's NLML_HPE_Test.py, B200-backed.

The reference driver (/root/reference/NLML_HPE_Test.py:182-449) loads two yaml files, the scripted
combined model, then loops images one at a time: FaceMesh -> x[1,1404] -> model(x) ->
round(np.degrees(t.item()), 3) (:271-273, :326-328, :410-412), and finally prints MAE statistics.
On the hot path are the model load and the forward; image decoding, MediaPipe and the dataset
readers are not.  This module keeps the config loading and the
post-processing convention and evaluates a *batch of pre-extracted feature vectors* in one call.
"""
from __future__ import annotations

import argparse

import numpy as np
import torch

from .config import load_config
from .NLML_HPE_Model_Builder import load_combined_model


def predict_degrees(model, features, device="cuda"):
    """features [N,1404] float32 (numpy or tensor) -> list of (yaw, pitch, roll) tuples in degrees rounded
    to 3 decimals, exactly the tuples the reference appends to pred_angles_NLML_HPE (:273)."""
    x = torch.as_tensor(np.asarray(features, dtype=np.float32)) if not isinstance(features, torch.Tensor) else features
    keep = ~(x == 0).all(dim=1)  # "no face" sentinel rows are skipped (:257-260, :396-399)
    with torch.no_grad():
        yaw, pitch, roll = model(x[keep].to(device))
    ang = np.degrees(torch.cat([yaw, pitch, roll], 1).cpu().numpy().astype(np.float64))
    return [tuple(round(float(v), 3) for v in row) for row in ang], keep.cpu().numpy()


def compute_mae(pred, true):
    pred, true = np.asarray(pred, dtype=np.float64), np.asarray(true, dtype=np.float64)
    err = np.abs(pred - true)
    return err.mean(0), err.std(0)


def NLML_HPE_Tester(argv=None):
    parser = argparse.ArgumentParser(description="Evaluate the Encoder+heads model on pre-extracted features")
    parser.add_argument("--features_npz", required=True, help="npz with 'X' [N,1404] and optional 'angles' [N,3] (deg)")
    parser.add_argument("--model", default="models/combined_model_scripted.pth")
    args = parser.parse_args(argv)
    device = "cuda"
    bins = load_config("configs/config_EncoderTrainer.yaml")      # :188-197
    cfg = load_config("configs/config_NLML_HPE_Test.yaml")        # :200-207 (tolerates the '=' line)
    model = load_combined_model(args.model)
    data = np.load(args.features_npz)
    pred, keep = predict_degrees(model, data["X"], device)
    print(f"processed {keep.sum()} samples, {(~keep).sum()} without landmarks; val_set = {cfg.get('val_set')}")
    if "angles" in data:
        true = data["angles"][keep]
        lo = np.array([bins[k]["min_bin"] for k in ("yaw_bins", "pitch_bins", "roll_bins")])
        hi = np.array([bins[k]["max_bin"] for k in ("yaw_bins", "pitch_bins", "roll_bins")])
        inside = ((true >= lo) & (true <= hi)).all(1)            # range filter of :245-247
        mae, std = compute_mae(np.asarray(pred)[inside], true[inside])
        print("MAE  yaw/pitch/roll/mean = %.3f %.3f %.3f %.3f" % (*mae, mae.mean()))
        print("STD  yaw/pitch/roll      = %.3f %.3f %.3f" % tuple(std))
    return pred


if __name__ == "__main__":
    NLML_HPE_Tester()
