"""Throughput of the bit-exact Powell fit (nlml_tucker_powell_f64). Usage: python scripts/time_powell.py [n]"""
import sys; sys.path.insert(0, ".")
import numpy as np, torch, bench
from nlml_hpe_b200 import synthetic
from nlml_hpe_b200.tucker import TuckerFitter
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 8
art, rows = bench.load_artifacts()
X = synthetic.make_features_torch(n, art["W"], *rows, U_id=art["U_id"], seed=1234, device="cuda")
fit = TuckerFitter(art["W"], *rows, device="cuda:0")
fit.powell(X[:148]); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); P, fun, nfev = fit.powell(X, return_info=True); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"powell n={n}: {ms:.1f} ms -> {n / ms * 1e3:.0f} poses/s; mean nfev {nfev.float().mean().item():.0f}; "
      f"{ms * 1e-3 * 1.965e9 * 148 / (nfev.sum().item()):.0f} SM-cycles per evaluation", flush=True)
