"""Extended-precision ground truth for the preconditioner (ORACLE -- test infrastructure only).

The reference holds no golden ApplyInverse vectors, and two correct FP64 implementations of
the same algorithm (sparse LU solves in the oracle, explicit inverses on the GPU) differ by cond(A11) * eps.
To decide which differences are rounding and which are bugs, this module runs the *same* restated algorithm
(oracle/hymls.py: src/HYMLS_Preconditioner.cpp:400-517,930-1070 and everything below it) in x87 extended
precision (np.longdouble, 64-bit mantissa, eps = 1.08e-19): every sparse/dense factorisation is an FP64 LU
followed by iterative refinement with residuals and updates in extended precision, every product, Householder
transform and assembly runs on longdouble arrays.  The result is ~2000x closer to the exact-arithmetic
preconditioner than any FP64 evaluation, which is enough to rank FP64 implementations:

    err(impl) = || x_impl - x_ext || / || x_ext ||

tests/test_gpu_parity.py asserts err(GPU) <= 2 err(oracle FP64) + 1e-14 and derives its tolerance from that.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import scipy.linalg as sla

from . import hymls as oh

LD = np.longdouble
REFINE_STEPS = 4


class _RefinedSparseLU:
    def __init__(self, M):
        self.A = sp.csr_matrix(M, dtype=LD)
        self.At = sp.csr_matrix(self.A.T)
        self.lu = spla.splu(sp.csc_matrix(self.A, dtype=np.float64), permc_spec="COLAMD")

    def solve(self, B, trans="N"):
        B = np.asarray(B, dtype=LD)
        A = self.A if trans == "N" else self.At
        x = self.lu.solve(np.asarray(B, dtype=np.float64), trans=trans).astype(LD)
        for _ in range(REFINE_STEPS):
            r = B - A @ x
            x = x + self.lu.solve(np.asarray(r, dtype=np.float64), trans=trans).astype(LD)
        return x


class _RefinedDenseLU:
    def __init__(self, M):
        self.A = np.asarray(M, dtype=LD)
        self.f = sla.lu_factor(np.asarray(M, dtype=np.float64))

    def solve(self, B):
        B = np.asarray(B, dtype=LD)
        x = sla.lu_solve(self.f, np.asarray(B, dtype=np.float64)).astype(LD)
        for _ in range(REFINE_STEPS):
            r = B - self.A @ x
            x = x + sla.lu_solve(self.f, np.asarray(r, dtype=np.float64)).astype(LD)
        return x


class Preconditioner(oh.Preconditioner):
    """oracle.hymls.Preconditioner evaluated in extended precision (all levels: the next level is type(self))."""
    dtype = LD

    @staticmethod
    def factor_sparse(M):
        return _RefinedSparseLU(M)

    @staticmethod
    def factor_dense(M):
        return _RefinedDenseLU(M)


class InversePreconditioner(oh.Preconditioner):
    """FP64, but with explicit inverses of the subdomain / separator / coarse blocks applied by GEMV -- a CPU
    model of the arithmetic the GPU path performs (used to tell rounding of the formulation from kernel bugs)."""

    class _Inv:
        def __init__(self, M):
            self.F = np.linalg.inv(M.toarray() if sp.issparse(M) else np.asarray(M))

        def solve(self, B, trans="N"):
            return (self.F if trans == "N" else self.F.T) @ B

    @staticmethod
    def factor_sparse(M):
        return InversePreconditioner._Inv(M)

    @staticmethod
    def factor_dense(M):
        return InversePreconditioner._Inv(M)
