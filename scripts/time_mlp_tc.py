"""Per-phase cycle counts of the epilogue warps of linear_tc2_kernel, per launch (development aid, not part of the product).

Build the instrumented library and run on the GPU box:
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DNLML_MLP_TIMING \
       -o nlml_hpe_b200/libnlml_mlp_timing.so nlml_hpe_b200/csrc/tucker_fit.cu nlml_hpe_b200/csrc/mlp_forward.cu
  NLML_HPE_LIB=nlml_hpe_b200/libnlml_mlp_timing.so python scripts/time_mlp_tc.py
Every linear_tc2_kernel launch of one chunk (encoder layers 0-2, heads' first layer) writes, per CTA and epilogue warp, the
average cycles per tile spent in: bias staging, waiting for a partial accumulator, promotion (TMEM loads + adds),
activation / plane split / stores (the NLML_MT_STAMP marks in mlp_tc.cuh).  Written in round 1 after the GPU budget was
spent: the production instruction stream is unchanged by the instrumentation (checked), the script itself is untested.
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB, _lib, synthetic  # noqa: E402

art, rows = bench.load_artifacts()
model = MB.build_combined_model(*bench.state_dicts(art))
n = 148 * 128 * 8                                            # one chunk
X = synthetic.make_features_torch(n, art["W"], *rows, U_id=art["U_id"], seed=1, device="cuda")
model.predict(X)                                             # warm-up, workspaces
lib = _lib.load()
slices, sms, warps = 4, 148, 8
buf = torch.zeros((slices, sms, warps, 4), dtype=torch.float32, device="cuda")
lib.nlml_debug_mlp_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.nlml_debug_mlp_timing.restype = None
lib.nlml_debug_mlp_timing(buf.data_ptr(), slices)
model.predict(X)
torch.cuda.synchronize()
names = ["encoder layer 0 (1408 -> 1024)", "encoder layer 1 (1024 -> 512)", "encoder layer 2 (512 -> 256)", "heads layer 1 (128 -> 256) x 3"]
phases = ["bias staging", "wait for accumulator", "promotion", "activation + split + stores"]
t = buf.cpu().numpy()
for s in range(slices):
    used = t[s][t[s].sum(-1) > 0]
    print(names[s], f"({len(used)} warps reporting): total {used.sum(-1).mean():.0f} cycles per tile")
    for i, ph in enumerate(phases):
        print(f"   {ph:30s} {used[:, i].mean():9.0f}   (min {used[:, i].min():.0f}, max {used[:, i].max():.0f})")
