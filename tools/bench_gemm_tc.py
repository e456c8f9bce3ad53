"""Micro-benchmark of the tcgen05 tile GEMM test hook (run on the GPU box: python tools/bench_gemm_tc.py)."""
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np
from neural_inventory_control_b200 import _lib, _capi as K
lib=_lib.load()
dev='cuda:0'
for (M,N,Kd) in [(8192,512,512),(8192,512,192),(8192,64,512),(1024,512,512)]:
    A=torch.randn(M,Kd,device=dev); B=torch.randn(N,Kd,device=dev)/Kd**0.5; C=torch.zeros(M,N,device=dev); scr=torch.zeros(2*(M*Kd+N*Kd),device=dev)
    for n_pass in (3,1):
        def run(): 
            rc=lib.hdpo_debug_gemm_tc(A.data_ptr(),B.data_ptr(),C.data_ptr(),M,N,Kd,n_pass,scr.data_ptr(),None); assert rc==0, lib.hdpo_last_error()
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/20
        ref=(A.double()@B.double().T)
        err=((C.double()-ref).abs().max()/ref.abs().max()).item()
        print(f"{M}x{N}x{Kd} n_pass={n_pass}: {ms*1e3:.1f} us incl. split, {2*M*N*Kd/ms/1e9:.1f} TFLOP/s algorithmic, err {err:.2e}")
    # torch fp32 reference speed
    torch.backends.cuda.matmul.allow_tf32=False
    for _ in range(3): A@B.T
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(20): A@B.T
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/20
    print(f"   cuBLAS fp32 (library, for context): {ms*1e3:.1f} us, {2*M*N*Kd/ms/1e9:.1f} TFLOP/s")
