"""TEST INFRASTRUCTURE: compile the SIMT kernel sources for the host-thread emulator (cuda_emu.h).

Produces tests/emu/_build/libhdpo_emu.so exporting the same C ABI as libhdpo_b200.so, with "device"
pointers being host pointers. Only the CPU tests load it; the product package never does.
tcgen05/TMA kernels cannot be emulated and are excluded (they are compiled out under HDPO_EMU).
"""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "neural_inventory_control_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libhdpo_emu.so")
SOURCES = ["capi.cu", "step_kernels.cu", "rollout_small.cu", "rollout_small_unit.cu", "rollout_small_bwd_kq1.cu", "rollout_small_bwd_kq2.cu", "rollout_small_bwd_kq4.cu",
           "rollout_small_bwd_kq5.cu", "rollout_small_bwd_kq8.cu", "rollout_wide.cu", "rollout_sym.cu", "gemm_tc.cu", "rollout_api.cu", "philox.cu", "adam.cu"]


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(ROOT, "include"), HERE):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h", ".py")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    return h.hexdigest()


def build(verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, "stamp")
    digest = _digest()
    if os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    objs, procs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(OUT_DIR, s + ".o")
        cmd = ["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-pthread", "-DHDPO_EMU", "-x", "c++", "-I", HERE, "-I", CSRC,
               "-Wno-unknown-pragmas", "-ffp-contract=off", "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"g++ (emu) failed for {s}:\n{out}")
    subprocess.check_call(["g++", "-shared", "-pthread", "-o", LIB, *objs])
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
