"""B200-native core for the src/collision hot path of jectpro7/realtime-collision-detection.

Import as ``rcd_b200`` (see rcd_b200/__init__.py).  Layout:
  csrc/   hand-written sm_100a CUDA kernels + the C-ABI (include/rcd.h) -> librcd_b200.so
  host/   Python mirror of the reference's detector / spatial-index API over that C-ABI:
          host.spatial_index      <- src/collision/spatial_index.py      (SpatialIndex, SpatialPartitioner)
          host.collision_detection<- src/collision/collision_detection.py (CollisionDetector, CollisionPredictionModel)
          host.warning_system     <- src/collision/warning_system.py     (AlertManager classification)
          host.compute_node       <- src/compute/compute_node.py:20-321  (SpatialIndex, VehicleState, CollisionDetector)
          host.engine             whole-frame batch API (FrameEngine); host.slabs: multi-GPU slabs + halo
There is no CPU fallback: every query runs on the GPU through librcd_b200.so.
"""
__version__ = "0.1.0"
