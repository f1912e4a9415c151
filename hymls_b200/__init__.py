"""hymls_b200 -- B200-native implementation of the HYMLS preconditioner hot path.

Host-side mirror of the reference interface (HYMLS::Preconditioner, HYMLS::Solver) over the C ABI of
`libhymls_b200.so` (include/hymls_b200.h).  All numerical work happens in hand-written sm_100a CUDA
kernels inside that library; there is no CPU fallback and no dependence on the test oracle.
"""
from .api import (HymlsError, Preconditioner, Solver, lib_path, load_library, params_to_xml,  # noqa: F401
                  pid_map)
from . import galeri, io  # noqa: F401

__all__ = ["Preconditioner", "Solver", "HymlsError", "galeri", "io", "params_to_xml", "load_library", "lib_path",
           "pid_map"]
