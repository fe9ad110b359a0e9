"""Per-phase cycle counts of the run-time-rank kernel's four roles (development build:
scripts/build_dev.sh tune -DNLML_GEN_TUNE -DNLML_GEN_TIMING; NLML_HPE_LIB=build/dev/libnlml_tune.so).
Usage: python scripts/time_gen.py ri ry rp rr F T"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from nlml_hpe_b200 import synthetic, _lib
from nlml_hpe_b200.tucker import TuckerFitter
ranks = tuple(int(x) for x in sys.argv[1:5]); F = int(sys.argv[5]); T = int(sys.argv[6]); n = 148 * 128
G = synthetic.synthetic_core(ranks, F, seed=11, std=1.0)
rws = [synthetic.synthetic_cos_params(r, 20 + i) for i, r in enumerate(ranks[1:])]
Xb = torch.from_numpy(synthetic.make_features(2048, G, *rws, U_id=None, seed=4)).cuda().repeat(10, 1)[:n].contiguous()
fit = TuckerFitter(G, *rws, device="cuda:0")
fit.fit(Xb, 2, kernel="tensor_core_generic"); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fit.fit(Xb, T, kernel="tensor_core_generic"); e1.record(); torch.cuda.synchronize()
out = (ctypes.c_float * 64)()
_lib.load().nlml_debug_gen_timing(out)
t = np.array(out).reshape(4, 16)
env = {k: v for k, v in os.environ.items() if k.startswith("NLML_GEN_")}
print(f"{ranks} {env}: {e0.elapsed_time(e1) / T * 1e3:.1f} us/iter; cycles per iteration:")
names = [["uu_wait", "tile_wait", "dt_buf_wait", "issue_T", "gtile_wait", "ypr_wait", "dg_buf_wait", "issue_G"],
         ["features+UU", "bar1", "wait_D_T", "fold", "bar2", "step"],
         ["other", "wait_opbuf", "form", "fence+arrive", "wait_D_G", "promote", "du", "bars"],
         ["features", "linear"]]
for r, role in enumerate(("MMA", "T-reader", "G-former", "linear")):
    print("  " + role + ": " + ", ".join(f"{nm}={t[r, i]:.0f}" for i, nm in enumerate(names[r])) + f"  (sum {t[r].sum():.0f})")
