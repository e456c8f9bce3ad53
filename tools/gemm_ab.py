"""A/B of the layer GEMM forms (single-tile CTA pairs vs multi-tile pairs), timed alone with CUDA events.
usage: python tools/gemm_ab.py  (on the GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_inventory_control_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")


def time_gemm(M, N, K, n_pass, nbuf, reps=64):
    g = torch.Generator(device=dev).manual_seed(1)
    A = torch.randn(M, K, generator=g, device=dev)
    Bm = torch.randn(N, K, generator=g, device=dev) / K ** 0.5
    bias = torch.zeros(N, device=dev)
    sets = []
    for _ in range(nbuf):
        scratch = torch.empty(2 * (M * K + N * K), device=dev)
        Cm, Clo = torch.empty(M, N, device=dev), torch.empty(M, N, device=dev)
        rc = lib.hdpo_debug_gemm_tc(A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), M, N, K, n_pass, scratch.data_ptr(), None)
        assert rc == 0, lib.hdpo_last_error()
        sets.append((scratch, Cm, Clo))
    dbg = torch.zeros(8 * 4096, dtype=torch.int64, device=dev)

    def run(i):
        scratch, Cm, Clo = sets[i % nbuf]
        rc = lib.hdpo_debug_gemm_tc_timeline(A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), M, N, K, n_pass,
                                             scratch.data_ptr(), dbg.data_ptr(), None, 0, Clo.data_ptr(), bias.data_ptr())
        assert rc == 0, lib.hdpo_last_error()

    for i in range(nbuf):
        run(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us per launch


for (M, N, K) in [(8192, 512, 512), (4096, 512, 512), (2048, 512, 512), (8192, 512, 192), (8192, 512, 64), (4096, 512, 192)]:
    row = []
    for multi in (0, 1):
        lib.hdpo_debug_set_tc_multi(1 if multi else 0)
        for nbuf in (1, 16):
            row.append("%s/%s %.1f us" % ("multi" if multi else "single", "warm" if nbuf == 1 else "cold", time_gemm(M, N, K, 3, nbuf)))
    print(M, N, K, " | ".join(row), flush=True)
lib.hdpo_debug_set_tc_multi(-1)
