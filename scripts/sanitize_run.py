"""Small invocations of every product kernel family for compute-sanitizer (development aid; run once plain, then under the tool).
  python scripts/sanitize_run.py
  compute-sanitizer --tool memcheck python scripts/sanitize_run.py
Sizes are ragged on purpose (not multiples of 32 / 128 rows) and small enough for the tool's slowdown."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB, synthetic  # noqa: E402
from nlml_hpe_b200.tucker import TuckerFitter  # noqa: E402

art, rows = bench.load_artifacts()
model = MB.build_combined_model(*bench.state_dicts(art))
fit = TuckerFitter(art["W"], *rows, device="cuda:0")
X = synthetic.make_features_torch(4300, art["W"], *rows, U_id=art["U_id"], seed=3, device="cuda")
for n in (1, 77, 300, 4300):                      # Encoder+heads: TMA-store epilogues on ragged tiles
    y = model.predict(X[:n])
    assert torch.isfinite(y).all()
y = model.predict_host(X[:1000].cpu().numpy())
for kernel, n, T in (("tensor_core", 4300, 4), ("tensor_core", 130, 4), ("thread_per_sample", 300, 4), ("warp_per_sample", 40, 4),
                     ("cta_per_sample", 3, 4), ("tensor_core_generic", 200, 3)):
    P = fit.fit(X[:n], T, kernel=kernel)            # 4300 rows: projection GEMM + slab copy; 130: in-kernel phase A
    assert torch.isfinite(P).all(), kernel
P = fit.solve(X[:4300])
P = fit.solve(X[:100])
P = fit.powell(X[:2])
Ph = fit.fit_host(X[:700].cpu().numpy(), 3)
ranks = (8, 5, 5, 5)
G = synthetic.synthetic_core(ranks, 96, seed=11, std=1.0)
rws = [synthetic.synthetic_cos_params(r, 21 + i, base=art[f"optimized_{k}"][:3]) for i, (r, k) in enumerate(zip(ranks[1:], ("yaw", "pitch", "roll")))]
fg = TuckerFitter(G, *rws, device="cuda:0")
Xg = torch.from_numpy(synthetic.make_features(150, G, *rws, U_id=None, seed=5)).cuda()
P = fg.fit(Xg, 3)
torch.cuda.synchronize()
print("sanitize_run ok")
