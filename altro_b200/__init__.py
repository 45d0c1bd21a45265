"""altro_b200 -- B200-native batched AL-iLQR solve path (drop-in for bjack205/altro's hot path).

The product is the CUDA shared library altro_b200/libaltro_b200.so (C ABI in
include/altro_b200.h).  This package is a thin ctypes binding whose `BatchSolver` mirrors
`altro::ALTROSolver` (src/altro/altro_solver.hpp) method for method, for a batch of B problems.
There is no CPU fallback: if the library or a CUDA device is missing, construction raises.
"""
from .solver import (AltroB200Error, BatchSolver, ErrorCodes, SolveStatus, build_library,  # noqa: F401
                     default_options, device_count, load_library, solve_problem, make_solver)
from . import problems  # noqa: F401
