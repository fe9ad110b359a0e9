import sys, time, numpy as np, torch
sys.path.insert(0,'.')
import bench
from nlml_hpe_b200 import synthetic
from nlml_hpe_b200.tucker import TuckerFitter
art,rows=bench.load_artifacts()
n=1000000
X=synthetic.make_features_torch(n,art["W"],*rows,U_id=art["U_id"],seed=1,device="cuda")
Xh=torch.empty((n,1404),dtype=torch.float32,pin_memory=True); Xh.copy_(X); torch.cuda.synchronize()
fit=TuckerFitter(art["W"],*rows,device="cuda:0")
P=np.empty((n,8),np.float32)
fit.fit_host(Xh,3000,out=P)
ts=[]
for _ in range(3):
    t=time.perf_counter(); fit.fit_host(Xh,3000,out=P); ts.append(time.perf_counter()-t)
print('fit_host 1M: %.1f ms -> %.3f M poses/s'%(min(ts)*1e3, n/min(ts)/1e6))
D=fit.fit(X,3000).cpu().numpy()
print('host == device:', np.array_equal(D,P))
