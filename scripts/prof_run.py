"""Small single-purpose runs for ncu / timing on the GPU box (not part of the product).

  python scripts/prof_run.py tucker --n 37888 --iters 300 [--kernel thread_per_sample]
  python scripts/prof_run.py mlp --n 65536
  python scripts/prof_run.py solve --n 1000000
Prints CUDA-event timings; run once plain, then under ncu with the same arguments.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB  # noqa: E402
from nlml_hpe_b200 import synthetic  # noqa: E402
from nlml_hpe_b200.tucker import TuckerFitter  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["tucker", "mlp", "solve"])
    ap.add_argument("--n", type=int, default=37888)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--kernel", default="thread_per_sample")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    art, rows = bench.load_artifacts()
    X = synthetic.make_features_torch(a.n, art["W"], *rows, U_id=art["U_id"], seed=1, device="cuda")
    if a.what == "tucker":
        fit = TuckerFitter(art["W"], *rows, device="cuda:0")
        ms = timed(lambda: fit.fit(X, a.iters, kernel=a.kernel), a.reps)
        best = min(ms)
        print(f"tucker {a.kernel} n={a.n} T={a.iters}: ms={['%.3f' % m for m in ms]}  "
              f"{a.n / best * 1e3:.0f} poses/s  {a.n * a.iters / best * 1e3 / 1e9:.3f} G sample-iters/s")
    elif a.what == "solve":
        fit = TuckerFitter(art["W"], *rows, device="cuda:0")
        ms = timed(lambda: fit.solve(X), a.reps)
        best = min(ms)
        _, ev = fit.solve(X, return_evals=True)
        ev = ev.cpu().numpy()
        wmax = ev[: len(ev) // 32 * 32].reshape(-1, 32).max(1)
        print(f"solve n={a.n}: ms={['%.3f' % m for m in ms]}  {a.n / best * 1e3:.0f} poses/s  evals mean {ev.mean():.2f} "
              f"max {ev.max()}  per-warp max mean {wmax.mean():.2f}")
    else:
        model = MB.build_combined_model(*bench.state_dicts(art))
        ms = timed(lambda: model.predict(X), a.reps)
        best = min(ms)
        print(f"mlp n={a.n}: ms={['%.3f' % m for m in ms]}  {a.n / best * 1e3:.0f} poses/s "
              f"{a.n / best * 1e3 * bench.MLP_FLOP_PER_POSE / 1e12:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
