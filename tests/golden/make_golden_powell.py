"""More outputs of the reference's shipped default fit, TD_Tester.Test = scipy Powell (TD_Tester.py:162-199), for the
comparison with the converged GPU solve: 96 samples (rows 0..95 of the seed-1234 batch; rows 0..7 repeat
tucker_golden.npz's powell_shipped_deg) with the objective value the reference's own `objective` reports at Powell's
result.  Run once, in the build container (needs /root/reference):  python tests/golden/make_golden_powell.py
Parity with Powell stays "unpinned" (scipy is an unpinned third-party dependency of the reference);
these vectors pin what THIS container's scipy returns so that the deviation of the drop-in's Test() is measured, not guessed."""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as mg  # noqa: E402
from nlml_hpe_b200 import synthetic  # noqa: E402


def _powell_with_loss(args):
    import torch
    import warnings
    W, x, oy, op_, or_ = args
    torch.set_num_threads(1)
    TD_Tester, _ = mg._import_reference()
    warnings.filterwarnings("ignore")
    import scipy.optimize as so
    captured = {}
    orig = so.minimize

    def spy(*a, **k):                      # the reference discards result.fun / nfev: keep them for the fixture
        r = orig(*a, **k)
        captured["fun"], captured["nfev"], captured["x"] = float(r.fun), int(r.nfev), np.array(r.x)
        return r
    TD_Tester.minimize = spy               # TD_Tester does `from scipy.optimize import minimize`
    y, p, r, _ = TD_Tester.Test(W, torch.tensor(x), W.shape[0], oy, op_, or_, None, None, None, None)
    return np.array([y, p, r], dtype=np.float64), captured["fun"], captured["nfev"], captured["x"]


if __name__ == "__main__":
    art = mg._load_shipped()
    W = art["W"]
    oy, op_, or_ = (art[f"optimized_{k}"][0:3, :] for k in ("yaw", "pitch", "roll"))
    X = synthetic.make_features(1000, W, oy, op_, or_, U_id=art["U_id"], seed=1234)
    n = 96
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as pool:
        res = list(pool.map(_powell_with_loss, [(W, X[i], oy, op_, or_) for i in range(n)]))
    import scipy
    out = {"idx": np.arange(n), "deg": np.stack([r[0] for r in res]), "loss": np.array([r[1] for r in res]),
           "nfev": np.array([r[2] for r in res]), "p": np.stack([r[3] for r in res]),
           "scipy_version": np.array(scipy.__version__)}
    np.savez_compressed(f"{HERE}/powell_golden.npz", **out)
    print("powell_golden.npz written:", {k: getattr(v, "shape", v) for k, v in out.items()})
    print("nfev min/median/max", out["nfev"].min(), np.median(out["nfev"]), out["nfev"].max())
