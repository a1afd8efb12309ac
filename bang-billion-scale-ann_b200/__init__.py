"""bang_b200 — B200-native (sm_100a) batched greedy Vamana search behind BANG's own entry points.

Directory name is `bang-billion-scale-ann_b200/`; import it as `bang_b200` through the root shim
(`/bang_b200.py`).  Contents: csrc/ (CUDA kernels, C ABI, loader, C++ shim, CLI), the ctypes mirror
of the reference's `BANGSearch<T>` interface (api.py), file formats, and fixture tooling.
"""
from . import build, formats, recall  # noqa: F401

__all__ = ["build", "formats", "recall"]
