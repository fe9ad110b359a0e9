"""3xTF32 tcgen05 building block against float64: mode 0 = K-major B operand, mode 1 = MN-major B operand read from the
K-major image of its transpose (decides whether one tile copy can serve both GEMMs of the run-time-rank kernel)."""
import sys, ctypes; sys.path.insert(0,".")
import numpy as np, torch
from nlml_hpe_b200 import _lib
lib=_lib.load()
torch.manual_seed(0)
for mode in (0, 1):
    for K,N in ((8,16),(8,96),(16,224),(8,256),(32,64),(48,48),(64,144)):
        A=torch.randn(128,K,device="cuda"); B=torch.randn(N,K,device="cuda"); D=torch.zeros(128,N,device="cuda")
        rc=lib.nlml_debug_tf32_gemm_mode(A.data_ptr(),B.data_ptr(),K,N,D.data_ptr(),mode)
        if rc: print("rc",rc,lib.nlml_last_error()); continue
        ref=(A.double()@B.double().T)
        err=(D.double()-ref).abs().max().item(); scale=ref.abs().max().item()
        f32=((A@B.T).double()-ref).abs().max().item()
        print("mode %d K=%d N=%d max abs err %.3e (ref max %.2f; plain fp32 matmul err %.3e)"%(mode,K,N,err,scale,f32), flush=True)
