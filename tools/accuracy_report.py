"""Accuracy of the GPU path against the extended-precision ground truth (oracle/extended.py), next to the FP64
oracle's own accuracy.  Run on the GPU box; prints one line per case and writes gpurun_out/accuracy_report.json.

    python tools/accuracy_report.py [--stages]
"""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hymls_b200 as hb  # noqa: E402
from oracle import extended as ox, hymls as oh  # noqa: E402
from tests.common import make_params  # noqa: E402
from tests.test_gpu_parity import CASES, dictify  # noqa: E402

LD = np.longdouble


def relerr(a, t):
    return float(np.linalg.norm(np.asarray(a, dtype=LD) - t) / np.linalg.norm(t))


def main():
    out = []
    for (eqn, dim, nx, sx, levels, cx, extra, _tol) in CASES:
        p = make_params(eqn, dim, nx, sx, levels, cx, **extra)
        A = hb.galeri.create_matrix(eqn, dim, nx)
        if eqn == "Stokes-C":
            A = -A
        A = sp.csr_matrix(A)
        tv = hb.galeri.create_testvector(A)
        b = np.random.default_rng(1).uniform(-1, 1, A.shape[0])
        t0 = time.time()
        X = ox.Preconditioner(A, p.copy(), tv); X.initialize(); X.compute()
        O = oh.Preconditioner(A, p.copy(), tv); O.initialize(); O.compute()
        P = hb.Preconditioner(A, dictify(p), tv); P.Initialize(); P.Compute()
        xt, xo, xg = X.apply_inverse(b), O.apply_inverse(b), P.ApplyInverse(b)
        rec = {"case": [eqn, dim, nx, sx, levels, cx, extra], "err_gpu_true": relerr(xg, xt),
               "err_oracle_true": relerr(xo, xt), "gpu_vs_oracle": relerr(xg, np.asarray(xo, dtype=LD)),
               "seconds": time.time() - t0}
        if "--stages" in sys.argv and levels >= 1:
            # A11^-1: GPU blocks and LAPACK inverses against refined extended-precision inverses
            off = P.DebugArray("a11off").astype(np.int64)
            F = P.DebugArray("a11inv")
            introw = P.DebugArray("introw").astype(np.int64)
            pos, eg, en = 0, 0.0, 0.0
            for sd in range(O.hid.num_subdomains()):
                idx = O.sd_int[sd]
                k = len(idx)
                if k == 0:
                    continue
                rows_gpu = introw[pos:pos + k]
                pos += k
                where = {int(r): q for q, r in enumerate(O.int_rows[idx])}
                perm = np.array([where[int(r)] for r in rows_gpu])
                blk = O.A11[idx[0]:idx[-1] + 1, idx[0]:idx[-1] + 1].toarray()
                it = ox._RefinedDenseLU(blk).solve(np.eye(k))[np.ix_(perm, perm)]
                npad = (k + 7) // 8 * 8
                G = F[off[sd]:off[sd] + npad * npad].reshape(npad, npad)[:k, :k]
                eg = max(eg, relerr(G, it))
                en = max(en, relerr(np.linalg.inv(blk)[np.ix_(perm, perm)], it))
            rec["a11inv_err_gpu"], rec["a11inv_err_lapack"] = eg, en
            ptr = P.DebugArray("redptr").astype(np.int64)
            col = P.DebugArray("redcol").astype(np.int64)
            val = P.DebugArray("redval")
            R = sp.csr_matrix((val, col, ptr), shape=(len(ptr) - 1, len(ptr) - 1)).toarray()
            Rt = X.schur_prec.reduced.toarray()
            Ro = O.schur_prec.reduced.toarray()
            sc = float(np.abs(Rt).max())
            rec["red_err_gpu"] = float(np.abs(R - Rt).max() / sc)
            rec["red_err_oracle"] = float(np.abs(Ro - Rt).max() / sc)
            boff = P.DebugArray("blkoff").astype(np.int64)
            BF = P.DebugArray("blkinv")
            eg = en = 0.0
            for bi, rows in enumerate(O.schur_prec.blocks):
                k = len(rows)
                if k == 0:
                    continue
                npad = (k + 7) // 8 * 8
                Mt = X.schur_prec.matrix[rows, :][:, rows].toarray()
                it = ox._RefinedDenseLU(Mt).solve(np.eye(k))
                G = BF[boff[bi]:boff[bi] + npad * npad].reshape(npad, npad)[:k, :k]
                eg = max(eg, relerr(G, it))
                en = max(en, relerr(np.linalg.inv(O.schur_prec.matrix[rows, :][:, rows].toarray()), it))
            rec["blkinv_err_gpu"], rec["blkinv_err_lapack"] = eg, en
        print(json.dumps(rec), flush=True)
        out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "accuracy_report.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
