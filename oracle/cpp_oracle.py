"""ctypes front end of the C++/OpenMP CPU restatement, oracle/cpp/hymls_oracle.cpp (ORACLE -- test
infrastructure and CPU baseline only; the product never imports this).

The C++ side restates the numerics (subdomain sparse LU, Schur assembly with Householder transform and dropping,
separator-block LU, recursion, coarse solve, GMRES/CG); the index maps are handed in per level, either from
oracle/partitioner.py (`maps_from_python_oracle`, small sizes) or from the library's host partitioner
(`maps_from_library`, any size; bit-exact with the former: tests/test_host_maps.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "cpp", "libhymls_oracle.so")
_LIB = None


def build():
    """compiles oracle/cpp (g++ -fopenmp); called by __graft_entry__.build()"""
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "cpp")], stdout=subprocess.DEVNULL)
    return _SO


def load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            build()
        lib = C.CDLL(_SO)
        vp, i64 = C.c_void_p, C.c_int64
        lib.ho_create.restype = vp
        lib.ho_create.argtypes = [C.c_int]
        lib.ho_destroy.argtypes = [vp]
        lib.ho_destroy.restype = None
        lib.ho_last_error.restype = C.c_char_p
        lib.ho_last_error.argtypes = [vp]
        lib.ho_set_matrix.argtypes = [vp, i64, vp, vp, vp, vp]
        lib.ho_set_partition.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]
        lib.ho_set_fix_gids.argtypes = [vp, C.c_int, vp]
        lib.ho_compute.argtypes = [vp, C.c_int]
        lib.ho_set_refinement.argtypes = [vp, C.c_int]
        lib.ho_apply_inverse.argtypes = [vp, vp, vp]
        lib.ho_solve.argtypes = [vp, C.c_int, vp, vp, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                 C.POINTER(C.c_int), C.POINTER(C.c_double), vp, C.c_int, C.POINTER(C.c_int)]
        lib.ho_stats.argtypes = [vp, vp]
        lib.ho_get_reduced.argtypes = [vp, C.c_int, C.POINTER(i64), C.POINTER(i64), vp, vp, vp]
        lib.ho_lu_fill.argtypes = [C.c_int, vp, vp, vp, C.c_int, vp, vp, C.POINTER(i64), C.POINTER(i64)]
        lib.ho_set_subdomain_ordering.argtypes = [vp, C.c_int]
        _LIB = lib
    return _LIB


def _flatten(interior, groups):
    """per-subdomain lists -> the flat arrays of ho_set_partition"""
    nsd = len(interior)
    int_ptr = np.zeros(nsd + 1, dtype=np.int64)
    sd_grp_ptr = np.zeros(nsd + 1, dtype=np.int64)
    for sd in range(nsd):
        int_ptr[sd + 1] = int_ptr[sd] + len(interior[sd])
        sd_grp_ptr[sd + 1] = sd_grp_ptr[sd] + len(groups[sd])
    int_gid = np.concatenate([np.asarray(i, dtype=np.int64) for i in interior]) if nsd else np.zeros(0, np.int64)
    flat = [g for sd in groups for g in sd]
    grp_ptr = np.zeros(len(flat) + 1, dtype=np.int64)
    for k, (_, nodes) in enumerate(flat):
        grp_ptr[k + 1] = grp_ptr[k] + len(nodes)
    grp_gid = (np.concatenate([np.asarray(n, dtype=np.int64) for _, n in flat]) if flat
               else np.zeros(0, dtype=np.int64))
    grp_type = np.asarray([t for t, _ in flat], dtype=np.int32)
    return nsd, int_ptr, np.ascontiguousarray(int_gid), sd_grp_ptr, grp_ptr, np.ascontiguousarray(grp_gid), grp_type


def maps_from_python_oracle(A, params, testvector=None):
    """[(interior, groups)] per level from oracle/partitioner.py (index sets only: identity matrices on the
    coarser levels, like tests/test_host_maps.py)"""
    from . import hymls as oh
    levels = max(params.sublist("Preconditioner").get("Number of Levels", 1), 1)
    lvl = oh.Preconditioner(A, params.copy(), testvector)
    lvl.initialize()
    out = [(lvl.hid.interior, lvl.hid.groups)]
    for l in range(levels - 1):
        s_ = lvl.schur_prec
        nv = len(s_.vsum_gids)
        lvl = oh.Preconditioner(sp.identity(nv, format="csr"), params.copy(), np.ones(nv), l + 1, s_.next_hid,
                                gids=s_.vsum_gids)
        lvl.initialize()
        out.append((lvl.hid.interior, lvl.hid.groups))
    return out


def maps_from_library(P):
    """[(interior, groups)] per level from an initialized hymls_b200.Preconditioner (host code, no GPU needed)"""
    out = []
    for l in range(P.NumLevels()):
        nsd = P.NumMySubdomains(l)
        out.append(([P.GetInteriorGroup(sd, l) for sd in range(nsd)], [P.GetSeparatorGroups(sd, l) for sd in range(nsd)]))
    return out


class Preconditioner:
    """HYMLS::Preconditioner(K, params, testVector) on the CPU (C++/OpenMP); same method names as oracle.hymls"""

    def __init__(self, A, params, testvector, maps, threads=0, refine_steps=0, fmatrix_ordering=True):
        self.lib = load()
        A = sp.csr_matrix(A)
        A.sort_indices()
        self.A = A
        self.n = A.shape[0]
        # BasePartitioner::SetParameters writes the defaults ("Fix GID 1" for Stokes, ...) into the list
        from .partitioner import CartesianPartitioner
        CartesianPartitioner(params, 0)
        prec = params.sublist("Preconditioner")
        self.levels = prec.get("Number of Levels", 1)
        if self.levels < 1:
            raise NotImplementedError("the C++ oracle covers Number of Levels >= 1 (oracle/hymls.py covers 0)")
        self.threads = threads
        self.h = C.c_void_p(self.lib.ho_create(self.levels))
        rp = np.ascontiguousarray(A.indptr, dtype=np.int64)
        ci = np.ascontiguousarray(A.indices, dtype=np.int32)
        v = np.ascontiguousarray(A.data, dtype=np.float64)
        tv = None if testvector is None else np.ascontiguousarray(testvector, dtype=np.float64)
        self._check(self.lib.ho_set_matrix(self.h, self.n, rp.ctypes.data, ci.ctypes.data, v.ctypes.data,
                                           None if tv is None else tv.ctypes.data))
        assert len(maps) >= self.levels
        for l in range(self.levels):
            nsd, ip, ig, sg, gp, gg, gt = _flatten(*maps[l])
            self._check(self.lib.ho_set_partition(self.h, l, nsd, ip.ctypes.data, ig.ctypes.data, sg.ctypes.data,
                                                  gp.ctypes.data, gg.ctypes.data, gt.ctypes.data))
        fix = []
        pos = 1
        while ("Fix GID %d" % pos) in prec:
            fix.append(prec["Fix GID %d" % pos])
            pos += 1
        fx = np.asarray(fix, dtype=np.int64)
        self._check(self.lib.ho_set_fix_gids(self.h, len(fix), fx.ctypes.data))
        # refine_steps > 0: every direct solve is refined with extended-precision residuals (checker only: makes the
        # oracle as accurate as the exact-arithmetic algorithm where plain FP64 LU solves are not)
        self.lib.ho_set_refinement(self.h, int(refine_steps))
        # subdomain solvers: the reference's F-matrix ordering + scaling + static pivots ("Custom Ordering" / "Custom
        # Scaling", default true in src/HYMLS_SparseDirectSolver.cpp:238-239), or a general minimum-degree ordering
        # with threshold partial pivoting
        self.lib.ho_set_subdomain_ordering(self.h, 1 if fmatrix_ordering else 0)

    def __del__(self):
        try:
            if self.h:
                self.lib.ho_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.ho_last_error(self.h).decode())

    def initialize(self):
        pass

    def compute(self):
        self._check(self.lib.ho_compute(self.h, self.threads))

    def apply_inverse(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros_like(b)
        self._check(self.lib.ho_apply_inverse(self.h, b.ctypes.data, x.ctypes.data))
        return x

    def solve(self, b, x0=None, method="GMRES", tol=1e-8, max_iters=1000, num_blocks=300, max_restarts=20):
        """right-preconditioned GMRES(m) / CG as oracle/krylov.py; returns x, iters, converged, history, seconds"""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        it, conv, hl = C.c_int(), C.c_int(), C.c_int()
        sec = C.c_double()
        hist = np.zeros(max_iters + 2)
        self._check(self.lib.ho_solve(self.h, 0 if method == "GMRES" else 1, b.ctypes.data, x.ctypes.data, tol,
                                      max_iters, num_blocks, max_restarts, C.byref(it), C.byref(conv), C.byref(sec),
                                      hist.ctypes.data, len(hist), C.byref(hl)))
        return x, it.value, bool(conv.value), hist[:hl.value].copy(), sec.value

    def stats(self):
        s = np.zeros(8)
        self.lib.ho_stats(self.h, s.ctypes.data)
        return {"compute_s": s[0], "factor_a11_s": s[1], "schur_s": s[2], "coarse_s": s[3], "nnz_factors": int(s[4]),
                "num_vsum": int(s[5]), "threads": int(s[6])}

    def reduced(self, level=0):
        n, nnz = C.c_int64(), C.c_int64()
        self._check(self.lib.ho_get_reduced(self.h, level, C.byref(n), C.byref(nnz), None, None, None))
        ptr = np.zeros(n.value + 1, dtype=np.int64)
        col = np.zeros(nnz.value, dtype=np.int32)
        val = np.zeros(nnz.value)
        self._check(self.lib.ho_get_reduced(self.h, level, C.byref(n), C.byref(nnz), ptr.ctypes.data, col.ctypes.data,
                                            val.ctypes.data))
        return sp.csr_matrix((val, col, ptr), shape=(n.value, n.value))


def lu_fill(A, b=None, fmatrix=False):
    """nnz(L) (unit diagonal included), nnz(U) of the oracle's sparse LU of A, and A^-1 b (or None).
    fmatrix: the reference's subdomain-solver ordering / scaling / static pivots (MatrixUtils::FillReducingOrdering)."""
    lib = load()
    Ac = sp.csc_matrix(A)
    Ac.sort_indices()
    n = Ac.shape[0]
    ap = np.ascontiguousarray(Ac.indptr, dtype=np.int32)
    ai = np.ascontiguousarray(Ac.indices, dtype=np.int32)
    ax = np.ascontiguousarray(Ac.data, dtype=np.float64)
    nl, nu = C.c_int64(), C.c_int64()
    x = None
    bp = xp = None
    if b is not None:
        bc = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(n)
        bp, xp = bc.ctypes.data, x.ctypes.data
    rc = lib.ho_lu_fill(n, ap.ctypes.data, ai.ctypes.data, ax.ctypes.data, 1 if fmatrix else 0, bp, xp, C.byref(nl),
                        C.byref(nu))
    if rc != 0:
        raise RuntimeError("ho_lu_fill failed (%d)" % rc)
    return nl.value, nu.value, x
