"""rfs-slam_b200: B200-native PHD measurement-update path of kykleung/RFS-SLAM.

The product is csrc/librfsb200.so (hand-written sm_100a CUDA behind the C ABI of
include/rfsb200.h).  The Python modules here are host plumbing for tests and the bench:

  capi   ctypes mirror of include/rfsb200.h
  phd    PHDUpdater: host-side mirror of RBPHDFilter::update() over the C ABI
  synth  deterministic synthetic 2-D range-bearing workloads
  dist   one-process-per-GPU sharding + the single weight all-reduce per step
"""
from . import capi, synth  # noqa: F401

__all__ = ["capi", "synth"]
