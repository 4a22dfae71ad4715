"""toss_b200 -- B200-native (sm_100a) population-fitness path of TOSS behind the reference's Python API.

The arithmetic lives in hand-written CUDA kernels (toss_b200/csrc) reached through the C ABI of
include/toss_b200.h; `toss_b200._native` is the ctypes binding.  There is no CPU fallback.
"""
__version__ = "0.1"
