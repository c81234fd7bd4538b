"""B200-native core for the src/collision hot path of jectpro7/realtime-collision-detection.

Import as ``rcd_b200`` (see rcd_b200/__init__.py).  Layout:
  csrc/   hand-written sm_100a CUDA kernels + the C-ABI (include/rcd.h) -> librcd_b200.so
  host/   Python mirror of the reference's detector / spatial-index API over that C-ABI
"""
__version__ = "0.1.0"
