"""restir_b200 -- B200-native (sm_100a) ReSTIR direct-illumination pipeline.

A drop-in for the frame hot path of HummaWhite/ReSTIR: C ABI in include/restir_b200.h, CUDA kernels and the
host scene layer in restir_b200/csrc, this package is the thin ctypes mirror used by tests and bench.py.
"""
from . import scenes  # noqa: F401
from .api import (  # noqa: F401
    REUSE_NONE, REUSE_SPATIAL, REUSE_SPATIOTEMPORAL, REUSE_TEMPORAL, TONEMAP_ACES, TONEMAP_FILMIC, TONEMAP_NONE,
    Camera, Denoiser, Frame, ReSTIRIndirect, RestirError, RstrParams, Scene, StripGroup, default_params, init, launch_count, lib, load_image, pinned_empty, pinned_free,
    write_jpg, write_png,
)
