import sys, ctypes; sys.path.insert(0,".")
import numpy as np, torch
from nlml_hpe_b200 import _lib
lib=_lib.load()
torch.manual_seed(0)
for K,N in ((8,16),(8,96),(16,224),(8,256),(32,64)):
    A=torch.randn(128,K,device="cuda"); B=torch.randn(N,K,device="cuda"); D=torch.zeros(128,N,device="cuda")
    rc=lib.nlml_debug_tf32_gemm(A.data_ptr(),B.data_ptr(),K,N,D.data_ptr())
    if rc: print("rc",rc,lib.nlml_last_error()); continue
    ref=(A.double()@B.double().T)
    err=(D.double()-ref).abs().max().item(); scale=ref.abs().max().item()
    f32=((A@B.T).double()-ref).abs().max().item()
    print("K=%d N=%d max abs err %.3e (ref max %.2f; plain fp32 matmul err %.3e)"%(K,N,err,scale,f32))
