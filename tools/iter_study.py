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
    ("cart sx8 L1 +tube", {"Separator Length": 8, "Number of Levels": 1, "Eliminate Tube Pressures With Velocities": True}),
    ("cart sx8 L2 cx4 +tube", {"Separator Length": 8, "Number of Levels": 2, "Coarsening Factor": 4, "Eliminate Tube Pressures With Velocities": True}),
    ("skew sx8 L1", {"Partitioner": "Skew Cartesian", "Separator Length": 8, "Number of Levels": 1}),
    ("skew sx8 L2 cx2", {"Partitioner": "Skew Cartesian", "Separator Length": 8, "Number of Levels": 2, "Coarsening Factor": 2}),
    ("skew sx8 L2 cx4", {"Partitioner": "Skew Cartesian", "Separator Length": 8, "Number of Levels": 2, "Coarsening Factor": 4}),
    ("skew sx4 L2 cx2", {"Partitioner": "Skew Cartesian", "Separator Length": 4, "Number of Levels": 2, "Coarsening Factor": 2}),
    ("skew sx4 L3 cx2", {"Partitioner": "Skew Cartesian", "Separator Length": 4, "Number of Levels": 3, "Coarsening Factor": 2}),
]:
    prec = dict(prec)
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
        print("%-22s its %4d conv %d res %.1e solve %.2fs setup %.1fs vsums %d nsd %d apply_bytes %.2f GB" % (
            desc, S.num_iter, S.info["converged"], S.info["explicit_rel_residual"], S.info["solve_seconds"], tc,
            st["num_vsum"], st["num_subdomains"], st["bytes_apply"] / 1e9), flush=True)
        del S, P
    except Exception as e:
        print(desc, "FAILED", str(e)[:200], flush=True)
