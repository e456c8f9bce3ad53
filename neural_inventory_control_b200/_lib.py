"""Loader of the product CUDA library (libhdpo_b200.so). No CPU fallback: fails loudly.

The library is built in-tree by `neural_inventory_control_b200.build` (nvcc, sm_100a). On a machine
with a CUDA device it is loaded with ctypes; when the .so is missing we try to build it once, and when
there is no CUDA device every compute entry point raises.
"""
import ctypes
import os

from . import _capi as K

ABI_VERSION = 2  # include/hdpo_b200.h HDPO_ABI_VERSION
_LIB = None
_DEVICE_OK = False  # cudaGetDeviceProperties costs ~2.5 ms: check once, not per call


class MissingExtension(RuntimeError):
    pass


def lib_path():
    # HDPO_LIB_PATH: another build of the SAME library (A/B timing of compile-time variants, tools/epi_ab.sh)
    return os.environ.get("HDPO_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libhdpo_b200.so")


def load(require_device=True):
    """Return the bound library. Raises MissingExtension when the .so cannot be loaded/built, and
    RuntimeError when `require_device` and no CUDA device is usable."""
    global _LIB, _DEVICE_OK
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            from . import build as _build
            try:
                _build.build()
            except Exception as e:  # noqa: BLE001
                raise MissingExtension(
                    f"libhdpo_b200.so is missing and could not be built ({e}). The HDPO engine has no CPU "
                    "fallback; run `python -m neural_inventory_control_b200.build`.") from e
        try:
            _LIB = K.bind(ctypes.CDLL(path))
        except OSError as e:
            raise MissingExtension(f"cannot load {path}: {e}") from e
        if _LIB.hdpo_abi_version() != ABI_VERSION:
            raise MissingExtension(f"{path}: ABI version {_LIB.hdpo_abi_version()} != {ABI_VERSION} (stale build?)")
    if require_device and not _DEVICE_OK:
        sm, major, minor = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        name = ctypes.create_string_buffer(128)
        rc = _LIB.hdpo_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor), name, 128)
        if rc != 0:
            raise RuntimeError("HDPO engine needs a CUDA device (sm_100a build, no CPU fallback): "
                               + _LIB.hdpo_last_error().decode())
        _DEVICE_OK = True
    return _LIB
