"""tracking_sdf_b200 — B200-native track + fuse hot path of mees/tracking_sdf.

The product is the C-ABI shared library built from csrc/ (include/tsdf_b200.h); this package
is only the thin ctypes mirror used by tests and bench.py.  There is no CPU fallback: if the
library is missing or no CUDA device is present, calls raise.
"""
from .capi import (Config, TrackStats, TsdfError, Tsdf, ShardGroup, load_library, library_path,
                   default_config, POINT_TO_PLANE, POINT_TO_POINT, HOST, DEVICE,
                   LAYOUT_REFERENCE, LAYOUT_XFASTEST)
from .api import SDF, CameraTracking

__all__ = ["Config", "TrackStats", "TsdfError", "Tsdf", "ShardGroup", "load_library", "library_path",
           "default_config", "SDF", "CameraTracking", "POINT_TO_PLANE", "POINT_TO_POINT", "HOST", "DEVICE",
           "LAYOUT_REFERENCE", "LAYOUT_XFASTEST"]
