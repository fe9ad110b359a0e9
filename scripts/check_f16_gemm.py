import sys, ctypes; sys.path.insert(0,".")
import numpy as np, torch
from nlml_hpe_b200 import _lib
lib=_lib.load()
torch.manual_seed(0)
for mode in (2, 3):
    for K,N in ((16,16),(16,224),(48,96),(64,256),(32,48)):
        A=torch.randn(128,K,device="cuda"); B=torch.randn(N,K,device="cuda")*100; D=torch.zeros(128,N,device="cuda")
        rc=lib.nlml_debug_tf32_gemm_mode(A.data_ptr(),B.data_ptr(),K,N,D.data_ptr(),mode)
        if rc: print("rc",rc,lib.nlml_last_error()); continue
        ref=(A.double()@B.double().T)
        err=(D.double()-ref).abs().max().item(); scale=ref.abs().max().item()
        f32=((A@B.T).double()-ref).abs().max().item()
        print("mode %d K=%d N=%d max abs err %.3e (ref max %.2f; plain fp32 matmul err %.3e)"%(mode,K,N,err,scale,f32), flush=True)
