"""GMRES iteration counts for variants of the preconditioner parameters (development study)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hymls_b200 as hb  # noqa: E402

nx = int(sys.argv[1])
A = -hb.galeri.create_matrix("Stokes-C", 3, nx)
tv = hb.galeri.create_testvector(A)
n = A.shape[0]
xex = np.random.default_rng(42).uniform(-1, 1, n)
b = A @ xex
for desc, prec in [
    ("sx8 L2 cx4", {"Separator Length": 8, "Number of Levels": 2, "Coarsening Factor": 4}),
    ("sx8 L1", {"Separator Length": 8, "Number of Levels": 1}),
    ("sx8 L2 cx4 retain2", {"Separator Length": 8, "Number of Levels": 2, "Coarsening Factor": 4, "Retain Nodes": 2}),
    ("sx8 L1 retain2", {"Separator Length": 8, "Number of Levels": 1, "Retain Nodes": 2}),
    ("sx4 L2 cx4", {"Separator Length": 4, "Number of Levels": 2, "Coarsening Factor": 4}),
    ("sx4 L3 cx2", {"Separator Length": 4, "Number of Levels": 3, "Coarsening Factor": 2}),
]:
    prec = dict(prec)
    prec["Eliminate Tube Pressures With Velocities"] = True
    params = {"Problem": {"Equations": "Stokes-C", "Dimension": 3, "nx": nx, "ny": nx, "nz": nx},
              "Preconditioner": prec,
              "Solver": {"Krylov Method": "GMRES", "Initial Vector": "Random",
                         "Iterative Solver": {"Maximum Iterations": 500, "Num Blocks": 500, "Maximum Restarts": 0,
                                              "Convergence Tolerance": 1e-8}}}
    try:
        t = time.time()
        P = hb.Preconditioner(A, params, tv)
        P.Initialize(); P.Compute()
        tc = time.time() - t
        S = hb.Solver(P)
        S.ApplyInverse(b, seed=43)
        st = P.Stats()
        print("%-22s its %4d conv %d res %.1e solve %.2fs setup %.1fs vsums %d apply_bytes %.2f GB" % (
            desc, S.num_iter, S.info["converged"], S.info["explicit_rel_residual"], S.info["solve_seconds"], tc,
            st["num_vsum"], st["bytes_apply"] / 1e9), flush=True)
        del S, P
    except Exception as e:
        print(desc, "FAILED", str(e)[:200], flush=True)
