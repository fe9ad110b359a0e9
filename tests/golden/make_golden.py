"""Generate the golden fixtures by running the REAL reference in the build container.

Run once, here (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes (all small, committed):
  shipped_artifacts.npz   repack of the reference's shipped data artefacts the hot path loads:
                          outputs/features/Trained_data.npz (W, optimized_*), Factor_Matrices.npz
                          (U_*), models/{yaw,pitch,roll}_network.pth (state-dict tensors).
                          Data, not source; CoreTensor is dropped (unused, TD_Inference.py:46).
  tucker_golden.npz       outputs of TD_Tester.optimize_with_sgd / objective_torch+autograd / Test
  mlp_golden.npz          outputs of CombinedAnglePredictionModel (eager + torch.jit.script)

Inputs are NOT stored: they are regenerated from seeds by nlml_hpe_b200.synthetic.
The reference modules import matplotlib / mediapipe at module scope without using them on
this path (TD_Tester.py:15, NLML_HPE_Model_Builder.py:20); empty stub modules stand in.
"""
import contextlib
import io
import os
import sys
import types
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from nlml_hpe_b200 import synthetic  # noqa: E402


def _import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "mediapipe", "cv2_stub"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import TD_Tester  # noqa
    import NLML_HPE_Model_Builder  # noqa
    return TD_Tester, NLML_HPE_Model_Builder


def _load_shipped():
    td = np.load(f"{REF}/outputs/features/Trained_data.npz")
    fm = np.load(f"{REF}/outputs/features/Factor_Matrices.npz")
    out = {"W": td["W"], "optimized_yaw": td["optimized_yaw"], "optimized_pitch": td["optimized_pitch"],
           "optimized_roll": td["optimized_roll"]}
    for k in ("U_id", "U_yaw", "U_pitch", "U_roll"):
        out[k] = fm[k]
    for head in ("yaw", "pitch", "roll"):
        sd = torch.load(f"{REF}/models/{head}_network.pth", map_location="cpu")
        for k, v in sd.items():
            out[f"{head}_network.{k}"] = v.numpy()
    return out


def _sgd_one(args):
    """Worker: the reference's optimize_with_sgd on one sample (TD_Tester.py:127-159)."""
    W, x, oy, op_, or_, iters = args
    torch.set_num_threads(1)
    TD_Tester, _ = _import_reference()
    t = lambda a: torch.tensor(a, dtype=torch.float32)  # noqa: E731  (as TD_Tester.py:170-174)
    with contextlib.redirect_stdout(io.StringIO()):
        p = TD_Tester.optimize_with_sgd(t(W), t(x), None, W.shape[0], t(oy), t(op_), t(or_),
                                        num_iterations=iters)
    return p.detach().numpy().copy()


def _powell_one(args):
    W, x, oy, op_, or_ = args
    TD_Tester, _ = _import_reference()
    import warnings
    warnings.filterwarnings("ignore")
    y, p, r, _ = TD_Tester.Test(W, torch.tensor(x), W.shape[0], oy, op_, or_, None, None, None, None)
    return np.array([y, p, r], dtype=np.float64)


def make_tucker(art):
    TD_Tester, _ = _import_reference()
    W = art["W"]
    oy, op_, or_ = (art[f"optimized_{k}"][0:3, :] for k in ("yaw", "pitch", "roll"))  # TD_Inference.py:56-57
    out = {}
    pool = ProcessPoolExecutor(max_workers=os.cpu_count())

    # (1) shipped W, full T=3000, 16 samples (seed 1234 = first 16 rows of the 1k bench batch)
    X = synthetic.make_features(1000, W, oy, op_, or_, U_id=art["U_id"], seed=1234)
    n_full = 16
    out["sgd3000_shipped_P"] = np.stack(list(pool.map(_sgd_one, [(W, X[i], oy, op_, or_, 3000) for i in range(n_full)])))
    out["sgd3000_shipped_idx"] = np.arange(n_full)
    # (2) shipped W, short T=200, 64 samples (rows 100..163)
    idx = np.arange(100, 164)
    out["sgd200_shipped_P"] = np.stack(list(pool.map(_sgd_one, [(W, X[i], oy, op_, or_, 200) for i in idx])))
    out["sgd200_shipped_idx"] = idx
    # (3) synthetic core of the configured rank (BASELINE.json config 2), T=3000, 8 samples
    G = synthetic.synthetic_core((5, 3, 3, 3), 1404, seed=7)
    Xg = synthetic.make_features(1000, G, oy, op_, or_, U_id=None, seed=4321)
    out["sgd3000_syncore_P"] = np.stack(list(pool.map(_sgd_one, [(G, Xg[i], oy, op_, or_, 3000) for i in range(8)])))
    out["sgd3000_syncore_idx"] = np.arange(8)
    # (4) noise-free, off-noise and zero ("no face", FeatureExtractor.py:105-106) edge inputs, T=500
    Xe = np.stack([synthetic.make_features(1, W, oy, op_, or_, U_id=art["U_id"], seed=5, noise=0.0)[0],
                   np.zeros(1404, np.float32),
                   X[0] * 10.0])
    out["sgd500_edge_X"] = Xe
    out["sgd500_edge_P"] = np.stack(list(pool.map(_sgd_one, [(W, Xe[i], oy, op_, or_, 500) for i in range(3)])))
    # (5) objective_torch + autograd at random parameter points (gradient known-answers)
    rng = np.random.default_rng(99)
    Pq = np.concatenate([rng.uniform(-0.8, 0.8, (32, 3)), rng.normal(0, 0.03, (32, 5))], 1).astype(np.float32)
    Pq[:, 3] += 0.0248
    losses, grads = [], []
    t = lambda a: torch.tensor(a, dtype=torch.float32)  # noqa: E731
    for i in range(32):
        p = torch.tensor(Pq[i], requires_grad=True)
        loss = TD_Tester.objective_torch(p, t(W), t(X[i]), t(oy), t(op_), t(or_))
        loss.backward()
        losses.append(loss.item())
        grads.append(p.grad.numpy().copy())
    out["grad_P"], out["grad_loss"], out["grad_G"] = Pq, np.array(losses, np.float32), np.stack(grads)
    # (6) shipped default: scipy Powell through Test() (parity unpinned; optimum only), 8 samples
    out["powell_shipped_deg"] = np.stack(list(pool.map(_powell_one, [(W, X[i], oy, op_, or_) for i in range(8)])))
    import scipy
    out["powell_scipy_version"] = np.array(scipy.__version__)
    out["torch_version"] = np.array(torch.__version__)
    pool.shutdown()
    np.savez_compressed(f"{HERE}/tucker_golden.npz", **out)
    print("tucker_golden.npz written")
    for k, v in out.items():
        print(" ", k, getattr(v, "shape", v))


def make_mlp(art):
    _, MB = _import_reference()
    oy, op_, or_ = (art[f"optimized_{k}"] for k in ("yaw", "pitch", "roll"))
    enc_sd = synthetic.synthetic_encoder_state_dict(art["W"], oy, op_, or_, U_id=art["U_id"], seed=0)
    encoder = MB.LandmarkEncoder(1404, [(1, 3)] * 3)
    encoder.load_state_dict({k: torch.tensor(v) for k, v in enc_sd.items()})
    heads = []
    for head in ("yaw", "pitch", "roll"):
        net = MB.AnglePredictionNetwork(3)
        net.load_state_dict(torch.load(f"{REF}/models/{head}_network.pth", map_location="cpu"))
        heads.append(net)
    model = MB.CombinedAnglePredictionModel(encoder, *heads).eval()
    scripted = torch.jit.script(model)
    X = synthetic.make_features(1000, art["W"], oy[0:3], op_[0:3], or_[0:3], U_id=art["U_id"], seed=1234)
    xt = torch.tensor(X)
    with torch.no_grad():
        eager = torch.cat(model(xt), 1).numpy()
        jit = torch.cat(scripted(xt), 1).numpy()
        # reference-style batch=1 loop with .item() (NLML_HPE_Test.py:326-328) on the first 64 rows
        b1 = np.array([[t.item() for t in scripted(xt[i:i + 1])] for i in range(64)], dtype=np.float32)
        lat = encoder.encoder(xt).numpy()
        # edge inputs: all-zero ("no face"), large magnitude, single row
        Xe = np.stack([np.zeros(1404, np.float32), X[0] * 50.0, -X[1]])
        edge = torch.cat(scripted(torch.tensor(Xe)), 1).numpy()
        # fp64 run of the same modules = error-budget reference for the stated tolerance
        m64 = MB.CombinedAnglePredictionModel(encoder, *heads).double()
        f64 = torch.cat(m64(xt.double()), 1).numpy()
    model.float()
    out = {"angles_eager": eager, "angles_jit": jit, "angles_b1": b1, "latent": lat, "edge_X": Xe,
           "edge_angles": edge, "angles_f64": f64, "torch_version": np.array(torch.__version__)}
    np.savez_compressed(f"{HERE}/mlp_golden.npz", **out)
    print("mlp_golden.npz written; max|eager-jit| =", np.abs(eager - jit).max(),
          " max|f32-f64| deg =", np.degrees(np.abs(eager - f64).max()),
          " angle range deg", np.degrees(eager.min(0)), np.degrees(eager.max(0)))


if __name__ == "__main__":
    art = _load_shipped()
    np.savez_compressed(f"{HERE}/shipped_artifacts.npz", **art)
    print("shipped_artifacts.npz written", os.path.getsize(f"{HERE}/shipped_artifacts.npz"))
    make_mlp(art)
    if "--mlp-only" not in sys.argv:
        make_tucker(art)
