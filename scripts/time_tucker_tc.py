"""Per-phase cycle counts of the tensor-core Tucker kernel (development aid, not part of the product).

Build the instrumented library and run on the GPU box:
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DNLML_TC_TIMING \
       -o build/dev/libnlml_timing.so nlml_hpe_b200/csrc/tucker_fit.cu nlml_hpe_b200/csrc/mlp_forward.cu
  NLML_HPE_LIB=build/dev/libnlml_timing.so python scripts/time_tucker_tc.py
Rows 0..7 of every CTA's output then hold the average cycles per iteration of lane 0 of warps 0..7 (warps 0-3 angle
role, 4-7 identity role; warp w runs on SM sub-partition w % 4) in the phases marked NLML_TSTAMP(i) in tucker_fit.cu.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nlml_hpe_b200 import synthetic  # noqa: E402
from nlml_hpe_b200.tucker import TuckerFitter  # noqa: E402

art, rows = bench.load_artifacts()
n = 148 * 128 * 2
X = synthetic.make_features_torch(n, art["W"], *rows, U_id=art["U_id"], seed=1, device="cuda")
fit = TuckerFitter(art["W"], *rows, device="cuda:0")
P = fit.fit(X, 1000, kernel="tensor_core").cpu().numpy().reshape(-1, 128, 8)
names = ["publish+issue(T)", "feat p,r (+V publish/issue)", "feat yaw + linear term", "GEMM wait", "readback+reduce+grad",
         "CTA barrier", "clip+step", "-"]
print(f"{'phase':34s}" + "".join(f"  warp{w}(r{w // 4})" for w in range(8)))
for i in range(7):
    print(f"{names[i]:34s}" + "".join(f"{P[:, w, i].mean():10.0f}" for w in range(8)))
print(f"{'total':34s}" + "".join(f"{P[:, w, :7].sum(1).mean():10.0f}" for w in range(8)))
