"""Build recipe for libhdpo_b200.so (the C-ABI library of include/hdpo_b200.h).

Plain nvcc, sm_100a only, in-tree output so the .so travels to the GPU box with the repo snapshot:

    python -m neural_inventory_control_b200.build [--force] [--verbose]

No torch headers are involved: the Python side binds the C ABI with ctypes and passes raw device
pointers + the current CUDA stream handle.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libhdpo_b200.so")
STAMP = os.path.join(PKG, ".libhdpo_b200.stamp")

SOURCES = ["capi.cu", "step_kernels.cu", "rollout_small.cu", "rollout_small_unit.cu", "rollout_small_bwd_kq1.cu", "rollout_small_bwd_kq2.cu", "rollout_small_bwd_kq4.cu",
           "rollout_small_bwd_kq5.cu", "rollout_small_bwd_kq8.cu", "rollout_wide.cu", "rollout_sym.cu", "gemm_tc.cu", "wide_persist.cu", "rollout_api.cu", "philox.cu", "adam.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-warn-spills",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _digest():
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    roots = [CSRC, os.path.join(os.path.dirname(PKG), "include")]
    for root in roots:
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    return h.hexdigest()


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def build(force=False, verbose=False, extra_flags=()):
    """Compile every .cu under csrc/ for sm_100a into libhdpo_b200.so (skipped when up to date)."""
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    objs = []
    build_dir = os.path.join(PKG, "build")
    os.makedirs(build_dir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    hdr = hashlib.sha256(" ".join([*NVCC_FLAGS, *extra_flags]).encode())
    for root in (CSRC, os.path.join(os.path.dirname(PKG), "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    hdr.update(name.encode() + f.read())
    stamps = []
    for src in sources():
        obj = os.path.join(build_dir, os.path.basename(src) + ".o")
        with open(src, "rb") as f:
            key = hashlib.sha256(hdr.hexdigest().encode() + f.read()).hexdigest()
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(obj + ".key") and open(obj + ".key").read() == key:
            continue  # object up to date (same source, headers and flags)
        stamps.append((obj + ".key", key))
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-I", CSRC, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed for {src}\n{out}\n")
        elif verbose or "warning" in out.lower() or "spill" in out.lower():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed (see above)")
    for path, key in stamps:
        with open(path, "w") as f:
            f.write(key)
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    subprocess.check_call(cmd)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
