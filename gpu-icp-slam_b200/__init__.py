"""gpu-icp-slam_b200: B200-native particle-filter SLAM engine behind the entry points of
michaelwillett/GPU-ICP-SLAM's src/kernel.h.

The product is the CUDA library `libpfslam.so` (C ABI in include/pfslam.h).  This package is the
host-side mirror of the reference interface on top of it (ctypes), plus the multi-GPU
orchestration.  There is no CPU fallback: using the engine without the built CUDA library raises.
"""
from .engine import (  # noqa: F401
    Config, FrameResult, ParticleFilter, PfslamError, Scene, Lidar, lib_path, load_library,
    particleFilterInit, particleFilter, particleFilterFree, getPCData,
    PATH_GRID2D, PATH_KD, SCORE_EXACT, SCORE_FILTERED, SCORE_TILED, QUIRKS_REFERENCE, QUIRK_Q1, RING_DEPTH,
)
from . import scans  # noqa: F401

__all__ = [
    "Config", "FrameResult", "ParticleFilter", "PfslamError", "Scene", "Lidar", "lib_path",
    "load_library", "particleFilterInit", "particleFilter", "particleFilterFree", "getPCData",
    "scans",
]
