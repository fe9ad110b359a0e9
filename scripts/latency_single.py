import sys, time, numpy as np, torch
sys.path.insert(0,'.')
import bench
from nlml_hpe_b200 import TD_Tester, synthetic, NLML_HPE_Model_Builder as MB
art,rows=bench.load_artifacts()
X=synthetic.make_features(64,art["W"],*rows,U_id=art["U_id"],seed=3)
W=art["W"]
TD_Tester.Test(W, torch.from_numpy(X[0]), 5, *rows, None,None,None,None)
t=time.perf_counter()
for i in range(64): TD_Tester.Test(W, torch.from_numpy(X[i]), 5, *rows, None,None,None,None)
print('Test() converged: %.3f ms per call'%((time.perf_counter()-t)/64*1e3))
TD_Tester.TEST_SOLVER="sgd"
TD_Tester.Test(W, torch.from_numpy(X[0]), 5, *rows, None,None,None,None)
t=time.perf_counter()
for i in range(16): TD_Tester.Test(W, torch.from_numpy(X[i]), 5, *rows, None,None,None,None)
print('Test() sgd-3000: %.3f ms per call'%((time.perf_counter()-t)/16*1e3))
m=MB.build_combined_model(*bench.state_dicts(art))
xg=torch.from_numpy(X).cuda()
with torch.no_grad():
    m(xg[:1]); torch.cuda.synchronize()
    t=time.perf_counter()
    for i in range(64):
        y,p,r=m(xg[i:i+1]); a=(y.item(),p.item(),r.item())
    print('model(x) batch-1 + 3x .item(): %.3f ms per call'%((time.perf_counter()-t)/64*1e3))
