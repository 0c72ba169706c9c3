"""psdr_cuda_b200 — B200-native hot path of psdr-cuda (renderC / renderD) behind a C ABI.

capi   ctypes binding of include/psdr_b200.h
build  nvcc recipes for the in-tree shared libraries
"""
__version__ = "0.1.0"
