"""Build the UNMODIFIED reference into baseline/_ref/ so that `bench.py --impl reference` can run it on the GPU box.

The reference (MatiasAlvo/Neural_inventory_control) is a flat directory of Python modules without setup.py /
pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing to build (outcome recorded in
DESIGN.md). What a pip install would have produced - the importable modules - is produced here instead by
BYTE-COMPILING the modules where they lie under /root/reference into baseline/_ref/<module>.bin (sourceless CPython
bytecode, same interpreter image on the GPU box; not named *.pyc because snapshot tools drop those as caches). No reference source enters the repository: baseline/_ref/ is git-ignored (and not
gpurun-ignored, so it travels). Run in the build container only (__graft_entry__.build() calls it).
"""
import os
import py_compile
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HDPO_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "baseline", "_ref")
MODULES = ("shared_imports", "data_handling", "environment", "neural_networks", "loss_functions", "trainer",
           "quantile_forecaster")


def try_pip():
    """The contract's install command; returns (ok, one-line outcome)."""
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
           "/opt/wheelhouse", "--target", OUT, REF]
    try:
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    except Exception as e:  # noqa: BLE001
        return False, f"pip not runnable: {e}"
    last = [ln for ln in p.stdout.strip().splitlines() if ln.strip()][-1:] or [""]
    return p.returncode == 0, last[0][:200]


def build(verbose=True):
    if not os.path.isdir(REF):
        return None  # GPU box: the prebuilt files are used as they are
    os.makedirs(OUT, exist_ok=True)
    ok, msg = try_pip()
    with open(os.path.join(OUT, "INSTALL_OUTCOME.txt"), "w") as f:
        f.write(f"pip install --target baseline/_ref {REF}: {'ok' if ok else 'failed'}: {msg}\n")
        f.write("modules byte-compiled from the reference tree instead (baseline/build_ref.py)\n")
    for m in MODULES:
        src = os.path.join(REF, m + ".py")
        if os.path.exists(src):
            py_compile.compile(src, cfile=os.path.join(OUT, m + ".bin"), dfile=f"<reference>/{m}.py", doraise=True,
                               optimize=0)
    if verbose:
        print(f"[build_ref] {len(MODULES)} reference modules -> {OUT} (pip: {'ok' if ok else 'failed'}: {msg})")
    return OUT


if __name__ == "__main__":
    build()
