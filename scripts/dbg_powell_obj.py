import sys, ctypes; sys.path.insert(0, ".")
import numpy as np, torch, bench
from nlml_hpe_b200 import synthetic, _lib
from nlml_hpe_b200.tucker import TuckerFitter
art, rows = bench.load_artifacts()
g = np.load("tests/golden/powell_golden.npz")
X = synthetic.make_features(1000, art["W"], *rows, U_id=art["U_id"], seed=1234)
host = ctypes.CDLL("tests/hostcheck/libhostcheck.so"); vp = ctypes.c_void_p
W2 = np.ascontiguousarray(art["W"].reshape(135, -1)); r = [np.ascontiguousarray(x, np.float64) for x in rows]
fit = TuckerFitter(art["W"], *rows, device="cuda:0"); lib = _lib.load()
rng = np.random.default_rng(0)
tot = bad = 0
for i in (5, 11, 0):
    pts = g["p"][i][None, :] + rng.normal(0, 1e-3, (2000, 8)); pts[0] = g["p"][i]; pts[1] = 0
    pts = np.ascontiguousarray(pts)
    vh = np.zeros(len(pts))
    host.hostcheck_powell_objective(vp(W2.ctypes.data), 5, 3, 3, 3, 1404, vp(r[0].ctypes.data), vp(r[1].ctypes.data), vp(r[2].ctypes.data),
                                    vp(X[i].ctypes.data), vp(pts.ctypes.data), len(pts), vp(vh.ctypes.data))
    xd = torch.from_numpy(X[i]).cuda(); pd = torch.from_numpy(pts).cuda(); vd = torch.zeros(len(pts), dtype=torch.float64, device="cuda")
    _lib.check(lib.nlml_debug_powell_objective(fit._h, xd.data_ptr(), pd.data_ptr(), len(pts), vd.data_ptr()))
    vd = vd.cpu().numpy()
    ne = vd != vh
    tot += len(pts); bad += ne.sum()
    print(i, "unequal", ne.sum(), "max rel diff", np.abs(vd - vh).max() / vh.max(), "golden f", g["loss"][i], vh[0], vd[0])
print("total", tot, "unequal", bad)
