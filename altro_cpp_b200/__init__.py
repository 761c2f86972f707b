"""altro_cpp_b200 — B200-native batched AL-iLQR (host-side Python binding of the C ABI).

The product is the CUDA library ``libaltro_b200.so`` (C ABI in include/altro_b200.h);
this package only marshals arrays to it.  There is no CPU fallback: importing works
anywhere, every compute call needs a CUDA device.
"""
from . import problems  # noqa: F401
from .capi import BatchSolver, MultiBatchSolver, Options, register_model, precompile_model, default_options, lib, SolverError, set_default_engine  # noqa: F401

__all__ = ["problems", "BatchSolver", "MultiBatchSolver", "register_model", "precompile_model", "Options", "default_options", "lib", "SolverError", "set_default_engine"]
