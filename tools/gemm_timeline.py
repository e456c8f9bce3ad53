"""GPU box: per-CTA clock64 timeline of the tcgen05 tile GEMM (setup / first data / mainloop / epilogue / store)."""
import ctypes as C
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import torch
import abi_driver as D

be = D.CudaBackend()
lib = be.lib
lib.hdpo_debug_gemm_tc_timeline.argtypes = [C.c_void_p] * 3 + [C.c_int32] * 4 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
lib.hdpo_debug_gemm_tc_timeline.restype = C.c_int
for (M, N, K) in [(2048, 512, 512), (8192, 512, 512), (2048, 512, 192), (2048, 64, 512)]:
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda") / K ** 0.5
    Cm = torch.zeros(M, N, device="cuda")
    scratch = torch.zeros(2 * (M * K + N * K), device="cuda")
    rc = lib.hdpo_debug_gemm_tc(A.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, K, 3, scratch.data_ptr(), be.stream)
    assert rc == 0
    bn = 128 if N % 128 == 0 else 64
    n_cta = (M // 128) * (N // bn)
    dbg = torch.zeros(8 * n_cta, dtype=torch.int64, device="cuda")
    Clo = torch.zeros(M, N, device="cuda"); bias = torch.randn(N, device="cuda")
    for n_pass, epi in ((3, 4), (3, 0), (1, 4)):
        for rep in range(2):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.hdpo_debug_gemm_tc_timeline(A.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, K, n_pass, scratch.data_ptr(), dbg.data_ptr(), be.stream, epi, Clo.data_ptr(), bias.data_ptr())
            e1.record(); torch.cuda.synchronize()
            assert rc == 0, lib.hdpo_last_error()
        t = dbg.cpu().numpy().reshape(n_cta, 8).astype(np.float64)
        d = t - t[:, :1]
        med = np.median(d, axis=0)
        names = ["entry", "setup done", "first tile landed", "last MMA issued", "accum ready", "epilogue math done", "stores done", "exit"]
        print(f"--- {M}x{N}x{K} n_pass={n_pass} epi={epi}: {n_cta} CTAs, kernel {e0.elapsed_time(e1)*1e3:.1f} us (event, incl. launch); median cycles since entry:")
        print("    " + ", ".join(f"{n} {int(v)}" for n, v in zip(names, med)))
