"""nanogi_b200 — B200-native (sm_100a) implementation of nanogi's `pt` / `ptdirect` hot path.

The product is `libnanogi_gpu.so` (hand-written CUDA behind the C ABI of include/nanogi_gpu.h) plus the
C++ front end (`nanogi` CLI, scene loader, film writers). This Python package is only binding glue for
tests and bench.py (`capi`) and the synthetic scene generators (`scenes`).
"""
from . import capi, scenes  # noqa: F401

__all__ = ["capi", "scenes"]
