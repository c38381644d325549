"""arkflow_b200 — B200 (sm_100a) implementation of ArkFlow's per-batch processor stage.

The package holds only the hot path: CUDA kernels + C ABI (csrc/, libarkflow_b200.so) and the
host-side mirror of the reference's Processor / Buffer plugin surface.  See DESIGN.md.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
