import sys, numpy as np, torch
sys.path.insert(0,'.')
import bench
from oracle import mlp_oracle
from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB, synthetic
art,rows=bench.load_artifacts(); sds=bench.state_dicts(art)
m=MB.build_combined_model(*sds)
DEG=180/np.pi
for seed,n in ((99,2048),(1234,1000),(7,32768)):
    X=synthetic.make_features(n,art["W"],*rows,U_id=art["U_id"],seed=seed)
    out=m.predict(torch.from_numpy(X).cuda()).cpu().numpy()
    ref=mlp_oracle.forward(*sds,X); ref64=mlp_oracle.forward(*sds,X,dtype=torch.float64)
    print(seed,n,'max err vs f32 oracle %.3e deg, vs f64 %.3e deg; oracle f32 vs f64 %.3e'%(np.abs(out-ref).max()*DEG,np.abs(out-ref64).max()*DEG,np.abs(ref-ref64).max()*DEG))
