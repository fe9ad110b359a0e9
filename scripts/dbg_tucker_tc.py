import sys; sys.path.insert(0,".")
import numpy as np, torch, bench
from nlml_hpe_b200 import synthetic
from nlml_hpe_b200.tucker import TuckerFitter
art,rows=bench.load_artifacts()
X=synthetic.make_features(300,art["W"],*rows,U_id=art["U_id"],seed=1234)
fit=TuckerFitter(art["W"],*rows,device="cuda:0"); x=torch.from_numpy(X).cuda()
np.set_printoptions(precision=6, suppress=False, linewidth=200)
for T in (1,2,3,5):
    a=fit.fit(x,T,kernel="thread_per_sample").cpu().numpy(); b=fit.fit(x,T,kernel="tensor_core").cpu().numpy()
    d=np.abs(a-b)
    print("T",T,"max per component",d.max(0), "rows with max", d.max(1).argmax())
    print(" fp32",a[0]); print(" tc  ",b[0])
