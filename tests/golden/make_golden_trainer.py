"""Outputs of the reference's offline cosine-curve trainer (TD_Trainer.py) for the tests of nlml_hpe_b200.TD_Trainer:
the Fourier initial guesses and the Powell-fitted rows for the shipped factor matrices (whose fitted rows ARE the shipped
optimized_{yaw,pitch,roll}) and for a synthetic set of factor matrices (other bin counts and ranks).
Run once in the build container (needs /root/reference):  python tests/golden/make_golden_trainer.py"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")
import TD_Trainer  # noqa: E402  (imports torch, numpy, scipy only)


def synthetic_factors(seed=5):
    """Three factor matrices shaped like a Tucker decomposition of a finer angle grid: 21 x 4, 17 x 5, 13 x 2."""
    rng = np.random.default_rng(seed)
    out = []
    for n_bins, rank, lim in ((21, 4, 50), (17, 5, 40), (13, 2, 30)):
        w = np.linspace(-lim, lim, n_bins)
        cols = []
        for j in range(rank):
            a, b, c, d = rng.uniform(0.2, 0.6), rng.uniform(0.6, 2.5), rng.uniform(-3, 3), rng.uniform(-0.2, 0.2)
            cols.append(a * np.cos(b * np.radians(w) + c) + d + rng.normal(0, 0.01, n_bins))
        out.append((np.stack(cols, 1), w))
    return out


if __name__ == "__main__":
    a = np.load(f"{HERE}/shipped_artifacts.npz")
    bins = {"yaw": np.arange(-50, 51, 10), "pitch": np.arange(-40, 41, 10), "roll": np.arange(-30, 31, 10)}   # TD_main.py:79-120
    shipped = [(a[f"U_{k}"], bins[k]) for k in ("yaw", "pitch", "roll")]
    out = {}
    for tag, sets in (("shipped", shipped), ("synth", synthetic_factors())):
        init = TD_Trainer.estimate_init_Fourier_Trans(*sets)
        with contextlib.redirect_stdout(io.StringIO()):
            fit = TD_Trainer.Train(*sets)
        for k, (U, w), i0, f in zip(("yaw", "pitch", "roll"), sets, init, fit):
            out[f"{tag}_{k}_U"], out[f"{tag}_{k}_w"] = np.asarray(U, np.float64), np.asarray(w, np.float64)
            out[f"{tag}_{k}_init"], out[f"{tag}_{k}_fit"] = i0, f
    np.savez_compressed(f"{HERE}/trainer_golden.npz", **out)
    print("trainer_golden.npz written")
    for k in ("yaw", "pitch", "roll"):
        print(k, "max |Train(shipped U) - shipped optimized|", np.abs(out[f"shipped_{k}_fit"] - a[f"optimized_{k}"]).max())
