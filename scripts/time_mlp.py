"""CUDA-event timing of the Encoder+heads forward on device-resident features (development aid).
  NLML_HPE_LIB=build/dev/libnlml_<name>.so python scripts/time_mlp.py [n_samples]
Prints ms per forward, poses/s and the maximum deviation (degrees) from the FP32 CUDA-core chain on the first 32 768 samples."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
art, rows = bench.load_artifacts()
model = MB.build_combined_model(*bench.state_dicts(art))
X = synthetic.make_features_torch(n, art["W"], *rows, U_id=art["U_id"], seed=1, device="cuda")
out = model.predict(X)
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        model.predict(X)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 3)
print(f"{os.environ.get('NLML_HPE_LIB', 'product')}: n={n}: {best:.3f} ms -> {n / best / 1e3:.2f} M poses/s")
a = model.predict(X[:200000]).cpu().numpy()
b = out[:200000].cpu().numpy()
print("repeatable across calls:", np.array_equal(a[:min(n, 200000)], b[:min(n, 200000)]))
