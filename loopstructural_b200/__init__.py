"""B200-native discrete implicit-interpolation engine for LoopStructural's FDI / P1 hot path.

Layout
  csrc/            hand-written sm_100a CUDA kernels and the C-ABI of ``include/lsb200.h``
  _native.py       ctypes binding of ``liblsb200.so`` (fails loudly when it is not built)
  engine.py        owner of one device handle
  interpolators.py host-side mirror of the reference's DiscreteInterpolator API
  plugin.py        registers the engine behind LoopStructural's interpolator factory
  partition.py     z-slab decomposition for multi-GPU solves
"""
from .engine import Engine, SolveResult, effective_stencil_weights  # noqa: F401

__version__ = "0.1.0"
