"""Staged GPU checks against the oracle (development tool; run on the B200 box via gpurun)."""
import os
import sys
import time
import traceback

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hymls_b200 as hb  # noqa: E402
from oracle import hymls as oh, krylov as ok  # noqa: E402
from tests.common import make_params  # noqa: E402


def dictify(p):
    return {k: (dictify(v) if isinstance(v, dict) else v) for k, v in p.items()}


def rel(a, b):
    d = np.linalg.norm(np.asarray(a) - np.asarray(b))
    n = np.linalg.norm(b)
    return d / n if n > 0 else d


def stage(eqn, dim, nx, sx, levels, cx=None, detail=True, solve=True, **extra):
    name = "%s %dD nx=%d sx=%d L=%d cx=%s %s" % (eqn, dim, nx, sx, levels, cx, extra or "")
    print("=" * 100 + "\n" + name, flush=True)
    p = make_params(eqn, dim, nx, sx, levels, cx, **extra)
    A = hb.galeri.create_matrix(eqn, dim, nx)
    if eqn == "Stokes-C":
        A = -A
    A = sp.csr_matrix(A)
    tv = hb.galeri.create_testvector(A)
    n = A.shape[0]
    t = time.time()
    O = oh.Preconditioner(A, p.copy(), tv)
    O.initialize()
    O.compute()
    print("oracle compute %.2fs" % (time.time() - t), flush=True)
    pd = dictify(p)
    pd["Solver"] = {"Krylov Method": "CG" if eqn == "Laplace" else "GMRES", "Initial Vector": "Zero",
                    "Left or Right Preconditioning": "Right",
                    "Iterative Solver": {"Maximum Iterations": 200, "Convergence Tolerance": 1e-8,
                                         "Maximum Restarts": 1}}
    t = time.time()
    P = hb.Preconditioner(A, pd, tv)
    P.Initialize()
    ti = time.time() - t
    t = time.time()
    P.Compute()
    print("gpu init %.2fs compute %.3fs" % (ti, time.time() - t), P.Stats(), flush=True)
    if detail and levels >= 1:
        # A11 inverses
        off = P.DebugArray("a11off").astype(np.int64)
        F = P.DebugArray("a11inv")
        introw = P.DebugArray("introw").astype(np.int64)   # library interior ordering (boundary nodes first)
        worst = 0
        pos = 0
        for sd in range(O.hid.num_subdomains()):
            idx = O.sd_int[sd]
            k = len(idx)
            if k == 0:
                continue
            where = {int(r): q for q, r in enumerate(O.int_rows[idx])}
            perm = np.array([where[int(r)] for r in introw[pos:pos + k]])
            pos += k
            npad = (k + 7) // 8 * 8
            inv = np.linalg.inv(O.A11[idx[0]:idx[-1] + 1, idx[0]:idx[-1] + 1].toarray())[np.ix_(perm, perm)]
            G = F[off[sd]:off[sd] + npad * npad].reshape(npad, npad)[:k, :k]
            worst = max(worst, rel(G, inv))
        print("A11 inverse max rel diff: %.3e" % worst)
        # reduced matrix
        S = O.schur_prec
        ptr = P.DebugArray("redptr").astype(np.int64)
        col = P.DebugArray("redcol").astype(np.int64)
        val = P.DebugArray("redval")
        R = sp.csr_matrix((val, col, ptr), shape=(len(ptr) - 1, len(ptr) - 1))
        Rref = S.reduced
        print("reduced Schur rel diff: %.3e  (nnz gpu %d oracle %d)" % (
            abs(R - Rref).max() / abs(Rref).max(), R.nnz, Rref.nnz))
        # separator blocks
        boff = P.DebugArray("blkoff").astype(np.int64)
        BF = P.DebugArray("blkinv")
        worst = 0
        for b, rows in enumerate(S.blocks):
            k = len(rows)
            if k == 0:
                continue
            npad = (k + 7) // 8 * 8
            inv = np.linalg.inv(S.matrix[rows, :][:, rows].toarray())
            G = BF[boff[b]:boff[b] + npad * npad].reshape(npad, npad)[:k, :k]
            worst = max(worst, rel(G, inv))
        print("separator block inverse max rel diff: %.3e (blocks %d)" % (worst, len(S.blocks)))
    rng = np.random.default_rng(1)
    worst = 0
    for k in range(3):
        b = rng.uniform(-1, 1, n)
        xg = P.ApplyInverse(b)
        xo = O.apply_inverse(b)
        worst = max(worst, rel(xg, xo))
    print("ApplyInverse max rel diff vs oracle: %.3e" % worst, flush=True)
    if solve:
        xex = rng.uniform(-1, 1, n)
        b = A @ xex
        S_ = hb.Solver(P)
        x = S_.ApplyInverse(b)
        if eqn == "Laplace":
            xo, its, conv, h = ok.cg(lambda v: A @ v, b, np.zeros(n), O.apply_inverse, tol=1e-8, max_iters=200)
        else:
            xo, its, conv, h = ok.gmres(lambda v: A @ v, b, np.zeros(n), O.apply_inverse, side="Right", tol=1e-8,
                                        max_iters=200, max_restarts=1)
        print("solve: gpu its %d conv %d res %.2e (%.3fs) | oracle its %d conv %s | x rel diff %.2e" % (
            S_.num_iter, S_.info["converged"], S_.info["explicit_rel_residual"], S_.info["solve_seconds"], its, conv,
            rel(x, xo)), flush=True)
        m = min(len(h), len(S_.history))
        print("history max rel dev: %.2e" % np.max(np.abs(np.asarray(h[:m]) - S_.history[:m]) / np.asarray(h[:m])))
    return P


if __name__ == "__main__":
    print(hb.load_library().hymls_b200_version())
    cases = [
        dict(eqn="Laplace", dim=2, nx=16, sx=4, levels=1),
        dict(eqn="Laplace", dim=2, nx=32, sx=4, levels=2),
        dict(eqn="Stokes-C", dim=2, nx=16, sx=8, levels=0),
        dict(eqn="Stokes-C", dim=2, nx=32, sx=4, levels=2),
        dict(eqn="Laplace", dim=3, nx=16, sx=4, levels=2),
        dict(eqn="Stokes-C", dim=3, nx=8, sx=4, levels=1, Eliminate_Tube_Pressures_With_Velocities=True),
        dict(eqn="Stokes-C", dim=3, nx=16, sx=4, levels=2, cx=2, Eliminate_Tube_Pressures_With_Velocities=True),
        dict(eqn="Stokes-C", dim=3, nx=16, sx=8, levels=1, Eliminate_Tube_Pressures_With_Velocities=True),
    ]
    if len(sys.argv) > 1:
        cases = [cases[int(a)] for a in sys.argv[1:]]
    for c in cases:
        try:
            stage(**c)
        except Exception:
            traceback.print_exc()
            print("STAGE FAILED", c, flush=True)
