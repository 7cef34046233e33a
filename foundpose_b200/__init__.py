"""foundpose_b200: B200-native (sm_100a) implementation of FoundPose's per-crop inference hot path.

The package mirrors the reference's `utils.*` modules for that path (see foundpose_b200/utils)
on top of hand-written CUDA kernels reached through the C ABI in include/foundpose_b200.h.
"""

__version__ = "0.1.0"
