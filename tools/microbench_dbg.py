import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hicom_b200 import ops
from tools.microbench_linear import timeit
dev="cuda"
for (M,N,K) in [(128,256,64),(128,256,1152),(18944,256,1152),(10368,3584,1152)]:
    A=torch.randn(M,K,device=dev).bfloat16(); W=(0.02*torch.randn(N,K,device=dev)).bfloat16()
    us=timeit(lambda: ops.linear(A,W,None,None,0,False,0))
    # raw C-ABI call loop without the torch custom-op layer
    import ctypes
    from hicom_b200 import _cabi
    lib=_cabi.load(); C=torch.empty(M,N,device=dev,dtype=torch.bfloat16)
    st=ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    def raw():
        lib.hicom_linear(ctypes.c_void_p(A.data_ptr()),K,ctypes.c_void_p(W.data_ptr()),K,None,None,0,ctypes.c_void_p(C.data_ptr()),N,M,N,K,0,1,1,M,0,0,st)
    us2=timeit(raw)
    print(f"dbg={os.environ.get('HICOM_TC_DBG','0')} M={M} N={N} K={K}: op {us:.1f} us, raw C-ABI {us2:.1f} us", flush=True)
