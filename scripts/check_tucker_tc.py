import sys; sys.path.insert(0,".")
import numpy as np, torch, bench
from nlml_hpe_b200 import synthetic
from nlml_hpe_b200.tucker import TuckerFitter
art,rows=bench.load_artifacts()
g=np.load("tests/golden/tucker_golden.npz")
X=synthetic.make_features(1000,art["W"],*rows,U_id=art["U_id"],seed=1234)
fit=TuckerFitter(art["W"],*rows,device="cuda:0"); x=torch.from_numpy(X).cuda()
for T in (1,10,100,3000):
    a=fit.fit(x,T,kernel="thread_per_sample").cpu().numpy(); b=fit.fit(x,T,kernel="tensor_core").cpu().numpy()
    d=np.degrees(np.abs(a[:,:3]-b[:,:3])).max(1)
    print("T=%d tc vs fp32 kernel: max %.3e deg, median %.3e, u max %.3e"%(T,d.max(),np.median(d),np.abs(a[:,3:]-b[:,3:]).max()))
ref=g["sgd3000_shipped_P"]; b=fit.fit(x,3000,kernel="tensor_core").cpu().numpy()
print("tc vs reference golden T=3000: %.3e deg"%np.degrees(np.abs(b[:16,:3]-ref[:,:3])).max())
