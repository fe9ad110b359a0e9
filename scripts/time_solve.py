"""Device-resident timing of the converged solve. Usage: python scripts/time_solve.py [n]"""
import sys; sys.path.insert(0, ".")
import numpy as np, torch, bench
from nlml_hpe_b200 import synthetic
from nlml_hpe_b200.tucker import TuckerFitter
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
art, rows = bench.load_artifacts()
X = synthetic.make_features_torch(n, art["W"], *rows, U_id=art["U_id"], seed=1234, device="cuda")
fit = TuckerFitter(art["W"], *rows, device="cuda:0")
P = fit.solve(X); torch.cuda.synchronize()
best = 1e9
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fit.solve(X, out=P); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"solve n={n}: {best:.2f} ms -> {n / best * 1e3:.0f} poses/s", flush=True)
Xs = X[:3000]
a = fit.solve(Xs).cpu().numpy()        # below the projection threshold: in-kernel phase A
b = fit.solve(X[:8192])[:3000].cpu().numpy()   # projected by the tensor-core GEMM
print("max |P(in-kernel phase A) - P(tensor-core phase A)| deg:", np.abs(a[:, :3] - b[:, :3]).max() * 180 / np.pi, flush=True)
