"""B200-native HDPO rollout engine behind the Neural_inventory_control Python API.

Submodules are imported lazily: `spec` / `_capi` are torch-free; everything that touches the GPU
library goes through `_lib.load()` and fails loudly when libhdpo_b200.so or a CUDA device is missing.
"""
__version__ = "0.1.0"
