"""Development check of the run-time-rank tensor-core Tucker kernel (kernel="tensor_core_generic") against the CPU oracle:
short iteration counts first (T = 1 checks the linear term and the projection, T = 2.. the GEMMs), then timings.
Usage: python scripts/check_gen.py [quick]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from nlml_hpe_b200 import synthetic
from nlml_hpe_b200.tucker import TuckerFitter
from oracle import tucker_oracle

DEG = 180 / np.pi


def case(ranks, F, n, Ts, seed=11, time_n=0, time_T=300):
    G = synthetic.synthetic_core(ranks, F, seed=seed, std=1.0)
    rws = [synthetic.synthetic_cos_params(r, 20 + i) for i, r in enumerate(ranks[1:])]
    X = synthetic.make_features(n, G, *rws, U_id=None, seed=3)
    t0 = time.time()
    fit = TuckerFitter(G, *rws, device="cuda:0")
    torch.cuda.synchronize()
    print(f"ranks {ranks} F={F}: plan {time.time() - t0:.2f} s", flush=True)
    xg = torch.from_numpy(X).cuda()
    for T in Ts:
        P = fit.fit(xg, T, kernel="tensor_core_generic").cpu().numpy()
        ref = tucker_oracle.sgd_batched(G, X, *rws, iters=T)
        da = np.abs(P[:, :3] - ref[:, :3]).max() * DEG
        du = np.abs(P[:, 3:] - ref[:, 3:]).max()
        print(f"   T={T:5d}: max angle diff {da:.3e} deg, max u diff {du:.3e} (|u| max {np.abs(ref[:, 3:]).max():.3f}, "
              f"angles max {np.abs(ref[:, :3]).max() * DEG:.2f} deg)", flush=True)
    if time_n:
        Xb = torch.from_numpy(synthetic.make_features(min(time_n, 4096), G, *rws, U_id=None, seed=4)).cuda()
        Xb = Xb.repeat((time_n + Xb.shape[0] - 1) // Xb.shape[0], 1)[:time_n].contiguous()
        fit.fit(Xb, 3, kernel="tensor_core_generic")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fit.fit(Xb, time_T, kernel="tensor_core_generic")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"   {time_n} samples x T={time_T}: {ms:.1f} ms -> {ms / time_T * 1e3:.1f} us/iteration, "
              f"{time_n / (ms * 3000 / time_T * 1e-3):.0f} poses/s at T=3000", flush=True)
    fit.close()


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    case((2, 2, 1, 3), 37, 9, [1, 2, 3, 50])
    case((4, 3, 2, 4), 50, 200, [1, 2, 3, 100])
    case((5, 3, 3, 3), 1404, 300, [1, 2, 10, 300], time_n=148 * 128, time_T=300)
    case((8, 5, 5, 5), 1404, 256, [1, 2, 10, 150], time_n=148 * 128, time_T=100)
    case((3, 2, 2, 7), 64, 130, [1, 2, 50])
    if not quick:
        case((16, 8, 8, 8), 96, 128, [1, 2, 5], time_n=148 * 128, time_T=5)
