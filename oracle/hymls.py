"""Preconditioner oracle (ORACLE -- test infrastructure only; never imported by the product).

CPU restatement (numpy/scipy) of the reference's hot path:
  * Preconditioner::{Initialize,Compute,ApplyInverse,ComputeBorder}
        src/HYMLS_Preconditioner.cpp:279-588,930-1070
  * MatrixBlock (A11/A12/A21/A22 + per-subdomain blocks)  src/HYMLS_MatrixBlock.cpp:74-385
  * SchurComplement::{Construct,Construct11,Construct22}   src/HYMLS_SchurComplement.cpp:88-306
  * SchurPreconditioner::{InitializeOT,CreateVSumMap,InitializeBlocks,
        AssembleTransformAndDrop,ConstructSCPart,ComputeNextLevel,ApplyInverse,
        ApplyOT,ApplyBlockDiagonal,UpdateVsumRhs, bordered variants}
        src/HYMLS_SchurPreconditioner.cpp:234-1093,1474-1619
  * Householder::{Apply,ApplyR,Construct,Apply(MV)}        src/HYMLS_Householder.cpp:38-163,353-363
  * RestrictedOT::Apply                                    src/HYMLS_RestrictedOT.hpp:21-36
  * CoarseSolver::{Compute,ApplyInverse, bordered}         src/HYMLS_CoarseSolver.cpp:131-323,454-564
  * MatrixUtils::{DropByValue,PutDirichlet}                src/HYMLS_MatrixUtils.cpp:1010-1309

The per-subdomain solver is a sparse LU (scipy SuperLU), standing in for
Ifpack_SparseContainer<KLU>; the arithmetic it performs (x = A11(sd)^-1 b) is the same.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import scipy.linalg as sla

from .partitioner import OverlappingPartitioner

SMALL = 1e-14  # HYMLS_SMALL_ENTRY, src/HYMLS_Macros.hpp:29


def sign(x):
    # src/HYMLS_Householder.cpp:15-18 -- note sign(0) == 0
    return -1.0 if x < 0 else (1.0 if x > 0 else 0.0)


# ---------------------------------------------------------------------------
# MatrixUtils
# ---------------------------------------------------------------------------
def drop_by_value(A, droptol=SMALL, typ="RelZeroDiag"):
    """src/HYMLS_MatrixUtils.cpp:1010-1194"""
    A = sp.csr_matrix(A)
    if droptol == 0.0:
        return A
    rel = typ in ("Relative", "RelDropDiag", "RelZeroDiag", "RelFullDiag")
    abs_diag = typ in ("RelDropDiag", "RelZeroDiag", "RelFullDiag", "AbsZeroDiag", "AbsFullDiag", "Absolute")
    zero_diag = typ in ("RelZeroDiag", "AbsZeroDiag")
    full_diag = typ in ("RelFullDiag", "AbsFullDiag")
    n = A.shape[0]
    diag = A.diagonal()
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    cols = A.indices
    vals = A.data
    is_diag = rows == cols
    scal = np.ones(len(vals))
    if rel:
        scal = np.maximum(np.abs(diag[rows]), np.abs(diag[cols]))
    if abs_diag:
        scal = np.where(is_diag, 1.0, scal)
    keep = (np.abs(vals) > scal * droptol) & (np.abs(vals) > droptol)
    new_vals = vals.copy()
    if full_diag:
        keep = keep & ~is_diag
    elif zero_diag:
        z = is_diag & ~keep
        new_vals[z] = 0.0
        keep = keep | z
    r, c, v = rows[keep], cols[keep], new_vals[keep]
    if full_diag:
        dv = np.where(np.abs(diag) > droptol, diag, 0.0)
        r = np.concatenate([np.arange(n), r])
        c = np.concatenate([np.arange(n), c])
        v = np.concatenate([dv, v])
    out = sp.csr_matrix((v, (r, c)), shape=A.shape)
    out.sort_indices()
    return out


def put_dirichlet(A, row):
    """src/HYMLS_MatrixUtils.cpp:1229-1309 (A is lil or csr; returns csr). Pattern is kept."""
    A = sp.csr_matrix(A).copy()
    s, e = A.indptr[row], A.indptr[row + 1]
    cols = A.indices[s:e].copy()
    A.data[s:e] = np.where(cols == row, 1.0, 0.0)
    for r in cols:
        if r != row:
            s2, e2 = A.indptr[r], A.indptr[r + 1]
            m = A.indices[s2:e2] == row
            A.data[s2:e2][m] = 0.0
    return A


# ---------------------------------------------------------------------------
# Householder
# ---------------------------------------------------------------------------
def householder_dense_rows(X, v):
    """Householder::Apply(SerialDenseMatrix, v): X = H X, src/HYMLS_Householder.cpp:38-80"""
    v = v * sign(v[0])
    nrmv = np.linalg.norm(v)
    v1 = v[0] + nrmv
    if abs(v1) < SMALL or nrmv < SMALL:
        return
    fac1 = 1.0 / (nrmv * v1)
    fac2 = nrmv * X[0, :] + v @ X
    fac = fac1 * fac2
    X0 = v1 * fac - X[0, :]
    X[1:, :] = np.outer(v[1:], fac) - X[1:, :]
    X[0, :] = X0


def householder_dense_cols(X, v):
    """Householder::ApplyR: X = X H', src/HYMLS_Householder.cpp:83-126"""
    householder_dense_rows(X.T, v)


def householder_w(v):
    """Householder::Construct: normalised reflector w (or None), src/HYMLS_Householder.cpp:128-163"""
    v = np.array(v, dtype=np.result_type(np.asarray(v).dtype, np.float64))
    nrm = np.linalg.norm(v)
    v = v * sign(v[0])
    v[0] += nrm
    nrm = np.linalg.norm(v)
    if nrm < SMALL:
        return None
    return v / nrm


# ---------------------------------------------------------------------------
# Coarse solver
# ---------------------------------------------------------------------------
class CoarseSolver:
    def __init__(self, matrix, gids, prec_params, level, factor_sparse=None):
        self.factor_sparse = factor_sparse or (lambda M: spla.splu(sp.csc_matrix(M)))
        self.matrix = sp.csr_matrix(matrix)
        self.gids = np.asarray(gids, dtype=np.int64)
        self.level = level
        self.fix_gid = []
        pos = 1
        while ("Fix GID %d" % pos) in prec_params:
            self.fix_gid.append(prec_params["Fix GID %d" % pos])
            pos += 1
        self.V = self.W = self.C = None
        self.lu = None

    def set_border(self, V, W, C):
        self.V, self.W, self.C = V, W, C

    def compute(self):
        n = self.matrix.shape[0]
        if n == 0:
            return
        S = drop_by_value(self.matrix, SMALL, "RelFullDiag")
        lid = {int(g): i for i, g in enumerate(self.gids)}
        for g in self.fix_gid:
            if g not in lid:
                raise RuntimeError("fix GID %d not in matrix row map" % g)
            S = put_dirichlet(S, lid[g])
        self.reduced = S
        if self.V is not None:
            # AugmentedMatrix [S V; W' C], src/HYMLS_CoarseSolver.cpp:200-224
            S = sp.bmat([[S, sp.csr_matrix(self.V)], [sp.csr_matrix(self.W.T), sp.csr_matrix(self.C)]]).tocsc()
        self.lu = self.factor_sparse(sp.csc_matrix(S))
        self._lid = lid

    def apply_inverse(self, X):
        if self.matrix.shape[0] == 0:
            return X.copy()
        rhs = np.array(X, copy=True)
        for g in self.fix_gid:
            l = self._lid.get(g, -1)
            if l > 0:  # sic: 'lid > 0', src/HYMLS_CoarseSolver.cpp:289
                rhs[l] = 0.0
        return self.lu.solve(rhs)

    def apply_inverse_bordered(self, X, T):
        """[Y;S] = [A V; W' C]^-1 [X;T], src/HYMLS_CoarseSolver.cpp:454-564"""
        n = self.matrix.shape[0]
        rhs = np.array(X, copy=True)
        # (no Fix-GID zeroing in the augmented solve, :497-509)
        full = np.concatenate([rhs, T], axis=0)
        sol = self.lu.solve(full)
        return sol[:n], sol[n:]


# ---------------------------------------------------------------------------
# Preconditioner (one level)
# ---------------------------------------------------------------------------
class _DenseLU:
    def __init__(self, M):
        self.f = sla.lu_factor(M)

    def solve(self, B):
        return sla.lu_solve(self.f, B)


class Preconditioner:
    # arithmetic hooks: oracle/extended.py overrides them to run the same algorithm in extended precision
    dtype = np.float64

    @staticmethod
    def factor_sparse(M):
        """per-subdomain / coarse sparse LU (Ifpack_SparseContainer<KLU>, Amesos)"""
        return spla.splu(sp.csc_matrix(M), permc_spec="COLAMD")

    @staticmethod
    def factor_dense(M):
        """separator-block LU (Ifpack_DenseContainer -> LAPACK dgetrf/dgetrs)"""
        return _DenseLU(M)

    def __init__(self, A, params, testvector=None, level=0, hid=None, gids=None):
        self.A = sp.csr_matrix(A, dtype=self.dtype)
        self.A.sort_indices()
        self.params = params
        self.prec_params = params.sublist("Preconditioner")
        self.level = level
        self.hid = hid
        n = self.A.shape[0]
        self.gids = np.arange(n, dtype=np.int64) if gids is None else np.asarray(gids, dtype=np.int64)
        self.testvector = testvector
        self.max_level = self.prec_params.get("Number of Levels", 1)
        self.V = self.W = self.C = None
        self.computed = False
        self.initialized = False

    # -- Initialize, src/HYMLS_Preconditioner.cpp:279-394 --------------------
    def initialize(self):
        if self.hid is None:
            self.hid = OverlappingPartitioner(self.params, self.level,
                                              base_gids=None if self.level == 0 else self.gids)
        hid = self.hid
        n = self.A.shape[0]
        gmax = int(self.gids.max()) + 1 if n else 0
        g2r = -np.ones(gmax, dtype=np.int64)
        g2r[self.gids] = np.arange(n)
        self.g2r = g2r
        self.int_gids = hid.interior_map()
        self.sep_gids = hid.separator_map()
        self.int_rows = g2r[self.int_gids]
        self.sep_rows = g2r[self.sep_gids]
        assert (self.int_rows >= 0).all() and (self.sep_rows >= 0).all()
        assert len(self.int_rows) + len(self.sep_rows) == n, "partition does not cover the map"
        self.nI, self.nS = len(self.int_rows), len(self.sep_rows)
        # position of every row in the interior / separator ordering
        self.pos_int = -np.ones(n, dtype=np.int64)
        self.pos_int[self.int_rows] = np.arange(self.nI)
        self.pos_sep = -np.ones(n, dtype=np.int64)
        self.pos_sep[self.sep_rows] = np.arange(self.nS)
        nsd = hid.num_subdomains()
        self.sd_int = []    # positions (in interior ordering) per sd
        self.sd_sep = []    # positions (in separator ordering) of ALL separator nodes around sd
        self.sd_grp_ptr = []
        off = 0
        for sd in range(nsd):
            k = len(hid.interior[sd])
            self.sd_int.append(np.arange(off, off + k))
            off += k
            nodes = [g for (_, nn) in hid.groups[sd] for g in nn]
            self.sd_sep.append(self.pos_sep[g2r[np.asarray(nodes, dtype=np.int64)]] if nodes
                               else np.zeros(0, dtype=np.int64))
            ptr = np.cumsum([0] + [len(nn) for (_, nn) in hid.groups[sd]])
            self.sd_grp_ptr.append(ptr)
        if self.testvector is None:
            self.testvector = np.ones(n)
        self.testvector = np.asarray(self.testvector, dtype=self.dtype)
        self.tv_sep = self.testvector[self.sep_rows]
        if self.level < self.max_level:
            self.schur_prec = SchurPreconditioner(self)
            self.schur_prec.initialize()
        self.initialized = True
        self.computed = False

    # -- Compute, src/HYMLS_Preconditioner.cpp:400-517 ------------------------
    def compute(self):
        if not self.initialized:
            self.initialize()
        self.compute_blocks()
        if self.level >= self.max_level:
            S = self.construct_schur()
            self.schur_prec = CoarseSolver(drop_by_value(S, SMALL), self.sep_gids,
                                           self.prec_params, self.level, self.factor_sparse)
        self.compute_border()
        self.schur_prec.compute()
        self.computed = True

    def compute_blocks(self):
        """MatrixBlock::Compute + ComputeSubdomainSolvers (src/HYMLS_MatrixBlock.cpp:74-292)"""
        A = self.A
        Ar = A[self.int_rows, :]
        As = A[self.sep_rows, :]
        self.A11 = sp.csr_matrix(Ar[:, self.int_rows])
        self.A12 = sp.csr_matrix(Ar[:, self.sep_rows])
        self.A21 = sp.csr_matrix(As[:, self.int_rows])
        self.A22 = sp.csr_matrix(As[:, self.sep_rows])
        # subdomain solvers (MatrixBlock::ComputeSubdomainSolvers :210-292)
        self.sd_lu = []
        A11c = self.A11.tocsr()
        for sd in range(self.hid.num_subdomains()):
            idx = self.sd_int[sd]
            if len(idx) == 0:
                self.sd_lu.append(None)
                continue
            blk = A11c[idx[0]:idx[-1] + 1, idx[0]:idx[-1] + 1].tocsc()
            self.sd_lu.append(self.factor_sparse(blk))

    def a11_solve(self, B, trans=False):
        """MatrixBlock::ApplyInverse, src/HYMLS_MatrixBlock.cpp:311-385"""
        X = np.zeros_like(B)
        for sd, lu in enumerate(self.sd_lu):
            if lu is None:
                continue
            idx = self.sd_int[sd]
            X[idx] = lu.solve(B[idx], trans="T" if trans else "N")
        return X

    # SchurComplement::Construct11 / Construct22, src/HYMLS_SchurComplement.cpp:131-306
    def construct22(self, sd):
        s = self.sd_sep[sd]
        return self.A22[s, :][:, s].toarray()

    def construct11(self, sd):
        s = self.sd_sep[sd]
        i = self.sd_int[sd]
        m = len(s)
        if len(i) == 0:
            return np.zeros((m, m))
        A12 = self.A12[i, :][:, s].toarray()
        B = self.sd_lu[sd].solve(A12)
        A21 = self.A21[s, :][:, i]
        return -(A21 @ B)

    def construct_schur(self):
        """SchurComplement::Construct, src/HYMLS_SchurComplement.cpp:88-129"""
        nS = self.nS
        rows, cols, vals = [], [], []
        for sd in range(self.hid.num_subdomains()):
            s = self.sd_sep[sd]
            Sk = self.construct11(sd)
            rr, cc = np.meshgrid(s, s, indexing="ij")
            rows.append(rr.ravel()); cols.append(cc.ravel()); vals.append(Sk.ravel())
            # structure from the A22 part
        S = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(nS, nS)) if rows else sp.csr_matrix((nS, nS))
        # pattern of S is that of the assembled A22 sub-blocks (InsertGlobalValues of Construct22);
        # SumInto only touches existing entries -- all Construct11 entries lie inside that pattern
        # for one-layer separators; keep the full sum here.
        return sp.csr_matrix(S + self.A22)

    # -- borders, src/HYMLS_Preconditioner.cpp:519-588,844-918 ---------------
    def set_border(self, V, W=None, C=None):
        if V is None:
            self.V = self.W = self.C = None
        else:
            V = np.asarray(V, dtype=self.dtype).reshape(self.A.shape[0], -1)
            self.V = V
            self.W = V if W is None else np.asarray(W, dtype=self.dtype).reshape(self.A.shape[0], -1)
            m = V.shape[1]
            self.C = np.zeros((m, m)) if C is None else np.asarray(C, dtype=self.dtype)
        self.computed = False

    def compute_border(self):
        if self.V is None:
            if hasattr(self.schur_prec, "set_border"):
                self.schur_prec.set_border(None, None, None)
            return
        self.V1, self.V2 = self.V[self.int_rows], self.V[self.sep_rows]
        self.W1, self.W2 = self.W[self.int_rows], self.W[self.sep_rows]
        self.Q1 = self.a11_solve(self.V1)
        sV = self.V2 - self.A21 @ self.Q1
        w1tmp = self.a11_solve(self.W1, trans=True)
        sW = self.W2 - self.A12.T @ w1tmp
        sC = self.C - self.W1.T @ self.Q1
        self.schur_prec.set_border(sV, sW, sC)

    # -- ApplyInverse, src/HYMLS_Preconditioner.cpp:930-1070 -----------------
    def _apply(self, B2, T):
        if not self.computed:
            raise RuntimeError("The preconditioner has not yet been computed.")
        b1 = B2[self.int_rows]
        b2 = B2[self.sep_rows]
        x1 = self.a11_solve(b1)
        y2 = self.A21 @ x1
        rhs = b2 - y2
        bordered = self.V is not None
        S = None
        if bordered:
            q = T - self.W1.T @ x1
            x2, S = self.schur_prec.apply_inverse_bordered(rhs, q)
        else:
            x2 = self.schur_prec.apply_inverse(rhs)
        y1 = self.A12 @ x2
        x1 = x1 - self.a11_solve(y1)
        if bordered:
            x1 = x1 - self.Q1 @ S
        X = np.zeros_like(B2)
        X[self.int_rows] += x1
        X[self.sep_rows] += x2
        return X, S

    def apply_inverse(self, B):
        """Epetra_Operator::ApplyInverse; with a border set: T = 0 and S is discarded (:594-605)"""
        B = np.asarray(B, dtype=self.dtype)
        B2 = B.reshape(self.A.shape[0], -1)
        T = None if self.V is None else np.zeros((self.V.shape[1], B2.shape[1]))
        X, _ = self._apply(B2, T)
        return X.reshape(B.shape)

    def apply_inverse_bordered(self, B, T):
        B2 = np.asarray(B, dtype=self.dtype).reshape(self.A.shape[0], -1)
        T2 = np.asarray(T, dtype=self.dtype).reshape(-1, B2.shape[1])
        if self.V is None:
            X, _ = self._apply(B2, None)
            return X, np.zeros_like(T2)
        return self._apply(B2, T2)


# ---------------------------------------------------------------------------
# SchurPreconditioner (one level)
# ---------------------------------------------------------------------------
class SchurPreconditioner:
    def __init__(self, prec):
        self.P = prec
        self.level = prec.level
        pp = prec.prec_params
        self.max_level = pp.get("Number of Levels", self.level)
        self.variant = pp.get("Preconditioner Variant", "Block Diagonal")
        self.dense_switch = pp.get("Dense Solvers on Level", 99)
        self.apply_dropping = pp.get("Apply Dropping", True)
        self.use_ot = pp.get("Apply Orthogonal Transformation", self.apply_dropping)
        if not (self.apply_dropping and self.use_ot and self.variant == "Block Diagonal"):
            raise NotImplementedError("oracle covers the default Block Diagonal / dropping / OT variant")
        self.V = self.W = self.C = None
        self.reduced_solver = None

    # Initialize, src/HYMLS_SchurPreconditioner.cpp:182-231
    def initialize(self):
        P, hid = self.P, self.P.hid
        tv = P.tv_sep
        nS = P.nS
        # InitializeOT :384-467 -- one sparse row w_g per LOCAL group, stored at the row of group[0]
        self.grp_pos = []   # per local group: positions in separator ordering
        self.grp_w = []     # reflector or None
        self.blocks = []    # per (sd, linked set of local groups): positions of non-Vsum nodes
        vsum = []
        for sd in range(hid.num_subdomains()):
            loc = hid.local_groups(sd)
            for gi in loc:
                nodes = np.asarray(hid.groups[sd][gi][1], dtype=np.int64)
                pos = P.pos_sep[P.g2r[nodes]]
                self.grp_pos.append(pos)
                self.grp_w.append(householder_w(tv[pos]))
                vsum.append(pos[0])
            for linked in hid.local_linked(sd):
                rows = []
                for gi in linked:
                    nodes = np.asarray(hid.groups[sd][gi][1], dtype=np.int64)
                    rows.extend(P.pos_sep[P.g2r[nodes[1:]]].tolist())
                self.blocks.append(np.asarray(rows, dtype=np.int64))
        self.vsum_pos = np.asarray(vsum, dtype=np.int64)      # CreateVSumMap :469-518
        self.vsum_gids = P.sep_gids[self.vsum_pos]
        self.next_hid = None
        if self.level + 1 < self.max_level:
            self.next_hid = hid.spawn_next_level(self.vsum_gids, self.vsum_gids)
        self.reduced_solver = None

    def set_border(self, V, W, C):
        self.V, self.W, self.C = V, W, C

    # ApplyOT :1236-1265 + Householder::Apply(MV) :353-363 : v <- 2 T'(T v) - v
    def apply_ot(self, v):
        out = -np.array(v, dtype=self.P.dtype, copy=True)
        v2 = np.asarray(v, dtype=self.P.dtype)
        for pos, w in zip(self.grp_pos, self.grp_w):
            if w is None:
                continue
            if v2.ndim == 1:
                out[pos] += 2.0 * w * (w @ v2[pos])
            else:
                out[pos] += 2.0 * np.outer(w, w @ v2[pos])
        return out

    # ConstructSCPart :877-986 (+ RestrictedOT::Apply)
    def _sc_part(self, sd, Sk):
        P, hid = self.P, self.P.hid
        s = P.sd_sep[sd]
        ptr = P.sd_grp_ptr[sd]
        v = P.tv_sep[s]
        ng = len(ptr) - 1
        for g in range(ng):
            a, b = ptr[g], ptr[g + 1]
            householder_dense_rows(Sk[a:b, :], v[a:b].copy())
            householder_dense_cols(Sk[:, a:b], v[a:b].copy())
        out = []
        first = ptr[:-1]
        out.append((s[first], Sk[np.ix_(first, first)]))
        for linked in hid.linked[sd]:
            loc = np.concatenate([np.arange(ptr[g] + 1, ptr[g + 1]) for g in linked]) \
                if linked else np.zeros(0, dtype=np.int64)
            loc = loc.astype(np.int64)
            out.append((s[loc], Sk[np.ix_(loc, loc)]))
        return out

    # AssembleTransformAndDrop :698-875
    def assemble(self):
        P, hid = self.P, self.P.hid
        nS = P.nS
        r1, c1, v1 = [], [], []
        r2, c2, v2 = [], [], []
        for sd in range(hid.num_subdomains()):
            for idx, blk in self._sc_part(sd, P.construct22(sd)):
                rr, cc = np.meshgrid(idx, idx, indexing="ij")
                r1.append(rr.ravel()); c1.append(cc.ravel()); v1.append(blk.ravel())
            for idx, blk in self._sc_part(sd, P.construct11(sd)):
                rr, cc = np.meshgrid(idx, idx, indexing="ij")
                r2.append(rr.ravel()); c2.append(cc.ravel()); v2.append(blk.ravel())
        r1 = np.concatenate(r1); c1 = np.concatenate(c1); v1 = np.concatenate(v1)
        r2 = np.concatenate(r2); c2 = np.concatenate(c2); v2 = np.concatenate(v2)
        # pass 1 = ReplaceGlobalValues (last writer wins), pass 2 = SumIntoGlobalValues
        key = r1 * nS + c1
        _, last = np.unique(key[::-1], return_index=True)
        last = len(key) - 1 - last
        M1 = sp.csr_matrix((v1[last], (r1[last], c1[last])), shape=(nS, nS))
        M2 = sp.csr_matrix((v2, (r2, c2)), shape=(nS, nS))
        # keep the full pattern (explicit zeros) like the FECrsMatrix does
        pat = sp.csr_matrix((np.ones(len(last)), (r1[last], c1[last])), shape=(nS, nS))
        M = (M1 + M2 + pat) - pat
        M = sp.csr_matrix(M)
        M.sort_indices()
        self.matrix = M

    # Compute :234-297
    def compute(self):
        P = self.P
        if P.nS == 0:
            return
        self.assemble()
        self.compute_next_level()
        M = self.matrix
        self.block_lu = []
        for rows in self.blocks:
            if len(rows) == 0:
                self.block_lu.append(None)
                continue
            blk = M[rows, :][:, rows].toarray()
            self.block_lu.append(self.P.factor_dense(blk))

    # ComputeNextLevel :520-629
    def compute_next_level(self):
        P = self.P
        vs = self.vsum_pos
        red = sp.csr_matrix(self.matrix[vs, :][:, vs])
        red = drop_by_value(red, SMALL, "RelDropDiag")
        self.reduced = red
        if self.level + 1 < self.max_level:
            ttv = self.apply_ot(P.tv_sep)
            next_tv = ttv[vs]
            nparams = P.params.copy()
            self.reduced_solver = type(P)(red, nparams, next_tv, self.level + 1,
                                                 self.next_hid, gids=self.vsum_gids)
            self.reduced_solver.initialize()
        else:
            self.reduced_solver = CoarseSolver(red, self.vsum_gids, P.prec_params, self.level + 1,
                                               P.factor_sparse)
        self.compute_border()
        self.reduced_solver.compute()

    # ComputeBorder :631-664
    def compute_border(self):
        if self.V is None:
            return
        self.bV = self.apply_ot(self.V)
        self.bW = self.apply_ot(self.W)
        vs = self.vsum_pos
        self.reduced_solver.set_border(self.bV[vs], self.bW[vs], self.C)

    def _block_diag(self, B):
        """ApplyBlockDiagonal :1311-1346 + UpdateVsumRhs :1435-1459"""
        Y = np.zeros_like(B)
        for rows, lu in zip(self.blocks, self.block_lu):
            if lu is None:
                continue
            Y[rows] = lu.solve(B[rows])
        Y[self.vsum_pos] = B[self.vsum_pos]
        return Y

    # ApplyInverse :1010-1093
    def apply_inverse(self, X):
        if self.P.nS == 0:
            return X.copy()
        if self.V is not None:
            m = self.V.shape[1]
            Y, _ = self.apply_inverse_bordered(X, np.zeros((m, X.shape[1]) if X.ndim > 1 else (m,)))
            return Y
        B = self.apply_ot(X)
        Y = self._block_diag(B)
        vs = self.vsum_pos
        Y[vs] = self.reduced_solver.apply_inverse(Y[vs])
        return self.apply_ot(Y)

    # bordered ApplyInverse :1517-1619
    def apply_inverse_bordered(self, X, T):
        B = self.apply_ot(X)
        Y = self._block_diag(B)
        vs = self.vsum_pos
        Y[vs] = 0.0                      # "note zeros in X2" (:1585)
        Tc = T - self.bW.T @ Y           # DenseUtils::MatMul(-1, borderW_, Y, 1, Tcopy)
        sol, S = self.reduced_solver.apply_inverse_bordered(B[vs], Tc)
        Y[vs] = sol
        return self.apply_ot(Y), S
