"""Determinism stress of the MLP chain: the same batch many times, every output compared bit for bit with the first."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB, synthetic  # noqa: E402

art, rows = bench.load_artifacts()
m = MB.build_combined_model(*bench.state_dicts(art))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 17161
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
X = synthetic.make_features_torch(n, art["W"], *rows, U_id=art["U_id"], seed=5, device="cuda")
ref = m.predict(X).clone()
lat = m.latent(X).clone()
bad = 0
for i in range(reps):
    out = m.predict(X)
    if not torch.equal(out, ref):
        d = (out != ref).any(1).nonzero().flatten()
        bad += 1
        if bad <= 5:
            print(f"rep {i}: {d.numel()} rows differ, first {d[:8].tolist()}, max |d| {float((out - ref).abs().max()):.3e}")
    if i % 50 == 0 and not torch.equal(m.latent(X), lat):
        print(f"rep {i}: latent differs")
print(f"NECK={os.environ.get('NLML_TC_NECK', '1')} TAIL={os.environ.get('NLML_TC_TAIL', '1')} n={n}: {bad} of {reps} runs differ")
