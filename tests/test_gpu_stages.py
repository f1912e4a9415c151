"""Stage-by-stage parity of Compute: the explicit subdomain inverses (row a7 of SURVEY 8a), the transformed +
dropped reduced Schur complement on the V-sums (a10-a12, a14) and the separator-block inverses (a13), read back
through the C ABI's test hook and compared with the EXTENDED-PRECISION evaluation of the restated reference
algorithm (oracle/extended.py).  Measured distances (profiles/r02_accuracy.md): <= 1.2e-15 for every stage on the
GPU (after the Newton-Schulz step), against 1e-13 .. 5e-12 for FP64 LAPACK / the FP64 oracle -- so the bound below
is a real bound, not slack.  ApplyInverse parity (test_gpu_parity.py) only sees the product of the stages."""
import numpy as np
import pytest
import scipy.sparse as sp

import hymls_b200 as hb
from oracle import extended as ox, hymls as oh
from tests.common import make_params
from tests.test_gpu_parity import dictify

pytestmark = pytest.mark.gpu

STAGE_TOL = 1e-13   # relative, against the extended-precision stage values
CASES = [
    ("Laplace", 2, 32, 4, 2, None, {}, STAGE_TOL),
    ("Stokes-C", 2, 32, 4, 2, None, {}, STAGE_TOL),
    ("Stokes-C", 3, 8, 4, 1, None, {"Partitioner": "Skew Cartesian"}, STAGE_TOL),
    ("Stokes-C", 3, 16, 4, 2, 2, {"Eliminate_Tube_Pressures_With_Velocities": True}, STAGE_TOL),
]


@pytest.mark.parametrize("eqn,dim,nx,sx,levels,cx,extra,tol", CASES)
def test_compute_stages_match_oracle(eqn, dim, nx, sx, levels, cx, extra, tol):
    p = make_params(eqn, dim, nx, sx, levels, cx, **extra)
    A = hb.galeri.create_matrix(eqn, dim, nx)
    if eqn == "Stokes-C":
        A = -A
    A = sp.csr_matrix(A)
    tv = hb.galeri.create_testvector(A)
    O = oh.Preconditioner(A, p.copy(), tv)
    O.initialize()
    O.compute()
    T = ox.Preconditioner(A, p.copy(), tv)      # ground truth: the same algorithm in extended precision
    T.initialize()
    T.compute()
    P = hb.Preconditioner(A, dictify(p), tv)
    P.Initialize()
    P.Compute()

    def dist(G, ref):
        ref = np.asarray(ref, dtype=np.longdouble)
        return float(np.linalg.norm(np.asarray(G, dtype=np.longdouble) - ref) / np.linalg.norm(ref))

    # (1) A11^-1 blocks.  The library orders the interior of a subdomain "nodes coupled to separators first"
    # (DESIGN.md 3); "introw" gives the matrix row of every interior position, the oracle keeps GID order.
    off = P.DebugArray("a11off").astype(np.int64)
    F = P.DebugArray("a11inv")
    introw = P.DebugArray("introw").astype(np.int64)
    assert sorted(introw.tolist()) == sorted(O.int_rows.tolist())
    dense11 = O.A11.toarray() if O.nI <= 4000 else None
    pos = 0
    for sd in range(O.hid.num_subdomains()):
        idx = O.sd_int[sd]
        k = len(idx)
        if k == 0:
            continue
        rows_gpu = introw[pos:pos + k]
        pos += k
        rows_ora = O.int_rows[idx]
        assert sorted(rows_gpu.tolist()) == sorted(rows_ora.tolist())      # same interior set per subdomain
        where = {int(r): q for q, r in enumerate(rows_ora)}
        perm = np.array([where[int(r)] for r in rows_gpu])                  # library position -> oracle position
        blk = (dense11[np.ix_(idx, idx)] if dense11 is not None
               else O.A11[idx[0]:idx[-1] + 1, idx[0]:idx[-1] + 1].toarray())
        inv = ox._RefinedDenseLU(blk).solve(np.eye(k))[np.ix_(perm, perm)]
        npad = (k + 7) // 8 * 8
        G = F[off[sd]:off[sd] + npad * npad].reshape(npad, npad)
        assert dist(G[:k, :k], inv) <= tol
        assert np.allclose(G[k:, k:], np.eye(npad - k), rtol=0, atol=1e-14)  # identity in the padding

    # (2) reduced Schur complement on the V-sums after transformation and dropping
    S = T.schur_prec
    ptr = P.DebugArray("redptr").astype(np.int64)
    col = P.DebugArray("redcol").astype(np.int64)
    val = P.DebugArray("redval")
    R = sp.csr_matrix((val, col, ptr), shape=(len(ptr) - 1, len(ptr) - 1))
    assert float(abs(R - S.reduced).max()) <= tol * float(abs(S.reduced).max())

    # (3) inverses of the non-V-sum separator blocks
    boff = P.DebugArray("blkoff").astype(np.int64)
    BF = P.DebugArray("blkinv")
    nb = 0
    for b, rows in enumerate(S.blocks):
        k = len(rows)
        if k == 0:
            continue
        npad = (k + 7) // 8 * 8
        inv = ox._RefinedDenseLU(S.matrix[rows, :][:, rows].toarray()).solve(np.eye(k))
        G = BF[boff[b]:boff[b] + npad * npad].reshape(npad, npad)[:k, :k]
        assert dist(G, inv) <= tol
        nb += 1
    assert nb > 0


@pytest.mark.parametrize("eqn,dim,nx,sx,levels,cx,extra,tol", [
    ("Stokes-C", 2, 32, 4, 2, None, {}, STAGE_TOL),
    ("Stokes-C", 3, 16, 4, 2, 2, {"Partitioner": "Skew Cartesian"}, STAGE_TOL),
])
def test_dense_schur_rows_path_matches_oracle(monkeypatch, eqn, dim, nx, sx, levels, cx, extra, tol):
    """The coarser levels form the rows of A21 A11^-1 with a DMMA GEMM per subdomain instead of the sparse
    accumulation (engine.cu: Level::schurGemm).  Forced on for every level here, so that the level-0 reduced
    Schur complement and separator blocks -- which the oracle can be compared with entry by entry -- go
    through it; ApplyInverse is compared as well."""
    monkeypatch.setenv("HYMLS_B200_SCHUR_GEMM", "1")
    test_compute_stages_match_oracle(eqn, dim, nx, sx, levels, cx, extra, tol)
    p = make_params(eqn, dim, nx, sx, levels, cx, **extra)
    A = sp.csr_matrix(-hb.galeri.create_matrix(eqn, dim, nx))
    tv = hb.galeri.create_testvector(A)
    O = oh.Preconditioner(A, p.copy(), tv)
    O.initialize()
    O.compute()
    P = hb.Preconditioner(A, dictify(p), tv)
    P.Initialize()
    P.Compute()
    b = np.random.default_rng(7).uniform(-1, 1, A.shape[0])
    x, xo = P.ApplyInverse(b), O.apply_inverse(b)
    assert np.linalg.norm(x - xo) <= 2e-11 * np.linalg.norm(xo)
