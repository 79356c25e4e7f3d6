"""jslpsolver_b200 -- B200-native (sm_100a) dense-tableau simplex + branch-and-cut behind
jsLPSolver's `solver.Solve()` JSON-model API.  The numerical path is hand-written CUDA in
libjslp_b200.so (C ABI: include/jslp_b200.h); there is no CPU fallback."""
from ._lib import JslpError, load as load_library  # noqa: F401
from .model import Model  # noqa: F401
from .solver import Solve, Solver, solver  # noqa: F401
from .tableau import DeviceContext, GpuTableau, default_context  # noqa: F401

__all__ = ["Solve", "Solver", "solver", "Model", "GpuTableau", "DeviceContext", "default_context",
           "JslpError", "load_library"]
