"""Development sweep of the run-time-rank kernel's configuration (needs scripts/build_dev.sh tune -DNLML_GEN_TUNE and
NLML_HPE_LIB=build/dev/libnlml_tune.so).  Usage: python scripts/sweep_gen.py ri ry rp rr F T [n]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from nlml_hpe_b200 import synthetic
from nlml_hpe_b200.tucker import TuckerFitter
ranks = tuple(int(x) for x in sys.argv[1:5]); F = int(sys.argv[5]); T = int(sys.argv[6])
n = int(sys.argv[7]) if len(sys.argv) > 7 else 148 * 128
G = synthetic.synthetic_core(ranks, F, seed=11, std=1.0)
rws = [synthetic.synthetic_cos_params(r, 20 + i) for i, r in enumerate(ranks[1:])]
Xb = torch.from_numpy(synthetic.make_features(min(n, 2048), G, *rws, U_id=None, seed=4)).cuda()
Xb = Xb.repeat((n + Xb.shape[0] - 1) // Xb.shape[0], 1)[:n].contiguous()
fit = TuckerFitter(G, *rws, device="cuda:0")
fit.fit(Xb, 2, kernel="tensor_core_generic"); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fit.fit(Xb, T, kernel="tensor_core_generic"); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
env = {k: v for k, v in os.environ.items() if k.startswith("NLML_GEN_")}
print(f"{ranks} {env}: {ms / T * 1e3:.1f} us/iter, {n / (ms * 3000 / T * 1e-3):.0f} poses/s at T=3000", flush=True)
