"""Stress of the MLP chain across plan creation / workspace growth: a fresh model per round, the size sequence of the
test file, every output compared bit for bit with round 0."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB, synthetic  # noqa: E402

art, rows = bench.load_artifacts()
sds = bench.state_dicts(art)
sizes = [1000, 1000, 3, 64, 1, 127, 128, 129, 255, 1000, 4097, 17161]
X = synthetic.make_features_torch(max(sizes), art["W"], *rows, U_id=art["U_id"], seed=5, device="cuda")
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ref = None
bad = 0
for r in range(rounds):
    m = MB.build_combined_model(*sds)
    outs = []
    for i, n in enumerate(sizes):
        if i == 1:
            outs.append(m.latent(X[:n]).clone())
        else:
            outs.append(m.predict(X[:n]).clone())
    torch.cuda.synchronize()
    if ref is None:
        ref = outs
        continue
    for i, (a, b) in enumerate(zip(outs, ref)):
        if not torch.equal(a, b):
            d = (a != b).any(1).nonzero().flatten()
            bad += 1
            print(f"round {r} size {sizes[i]}: {d.numel()} rows differ, first {d[:8].tolist()} max|d| {float((a - b).abs().max()):.3e}")
    m.invalidate()
print(f"{bad} mismatching outputs in {rounds} rounds")
