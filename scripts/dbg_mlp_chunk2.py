"""One process = one line: checksum of the GPU output and of the CPU oracle for the chunk-boundary test's data."""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle import mlp_oracle  # noqa: E402
from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB, synthetic  # noqa: E402

art, rows = bench.load_artifacts()
sds = bench.state_dicts(art)
X = synthetic.make_features(2048, art["W"], *rows, U_id=art["U_id"], seed=99)
ref = mlp_oracle.forward(*sds, X)
m = MB.build_combined_model(*sds)
for n in (1000, 3, 1, 127, 129):
    m.predict(torch.from_numpy(X[:n]).cuda())
big = torch.from_numpy(X).cuda().repeat(9, 1)[:16384 + 777].contiguous()
out = m.predict(big).cpu().numpy()
err = np.abs(out[:2048] - ref).max(1) * 180 / np.pi
print("gpu", hashlib.sha1(out.tobytes()).hexdigest()[:12], "cpu-oracle", hashlib.sha1(ref.tobytes()).hexdigest()[:12],
      "max err %.3e" % err.max(), "rows>1e-3:", np.nonzero(err > 1e-3)[0][:8], "threads", torch.get_num_threads())
