"""graphmat_b200 -- B200-native engine for GraphMat's hot path (send -> SpMSpV -> apply).

The product is the C-ABI library ``libgraphmat_b200.so`` (sources in ``csrc/``,
contract in ``/include/graphmat_b200.h``) plus the C++ drop-in headers in
``include/GraphMat``.  This Python package is only the ctypes binding the tests
and ``bench.py`` use; it has no CPU fallback: without the CUDA library every
entry point raises.
"""
from . import capi  # noqa: F401
from .capi import Graph, Vectors, lib, build_library  # noqa: F401
