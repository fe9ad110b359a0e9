import sys, numpy as np, torch
sys.path.insert(0,'.')
import bench
from oracle import mlp_oracle
from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB, synthetic
art,rows=bench.load_artifacts(); sds=bench.state_dicts(art)
X=synthetic.make_features(2048,art["W"],*rows,U_id=art["U_id"],seed=99)
ref=mlp_oracle.forward(*sds,X)
DEG=180/np.pi
for trial in range(6):
    m=MB.build_combined_model(*sds)
    # same call sequence as the test file
    m.predict(torch.from_numpy(X[:1000]).cuda()); m.latent(torch.from_numpy(X[:1000]).cuda())
    for n in (1,127,128,129,1000):
        m.predict(torch.from_numpy(X[:n]).cuda())
    big=torch.from_numpy(X).cuda().repeat(9,1)[:16384+777].contiguous()
    out=m.predict(big).cpu().numpy()
    full=np.tile(ref,(9,1))[:16384+777]
    err=np.abs(out-full).max(1)*DEG
    bad=np.nonzero(err>1e-3)[0]
    print('trial',trial,'max err %.3e'%err.max(),'bad rows',len(bad), bad[:10], (bad//128)[:10] if len(bad) else '')
    m.invalidate()
