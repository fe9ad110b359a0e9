"""Device-resident timing of the fixed-iteration Tucker fit on the shipped core. Usage: python scripts/time_fit.py [kernel] [n] [T]"""
import sys; sys.path.insert(0, ".")
import numpy as np, torch, bench
from nlml_hpe_b200 import synthetic
from nlml_hpe_b200.tucker import TuckerFitter
kernel = sys.argv[1] if len(sys.argv) > 1 else "tensor_core"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8 * 148 * 128
T = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
art, rows = bench.load_artifacts()
X = synthetic.make_features_torch(n, art["W"], *rows, U_id=art["U_id"], seed=1234, device="cuda")
fit = TuckerFitter(art["W"], *rows, device="cuda:0")
P = fit.fit(X, T, kernel=kernel); torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fit.fit(X, T, kernel=kernel, out=P); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"{kernel} n={n} T={T}: {best:.2f} ms -> {n / best * 1e3:.0f} poses/s, {best / T * 1e3 * 148 * 128 / n:.3f} us per iteration-wave", flush=True)
