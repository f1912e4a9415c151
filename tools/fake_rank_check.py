"""Runs the sharded code path of one rank on a single GPU (no NCCL: collectives are no-ops), for
compute-sanitizer.  Results are incomplete by construction; only memory safety is checked."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hymls_b200 as hb  # noqa: E402

nx, sx, rank, nranks = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
params = {"Problem": {"Equations": "Stokes-C", "Dimension": 3, "nx": nx, "ny": nx, "nz": nx},
          "Preconditioner": {"Separator Length": sx, "Number of Levels": 2, "Coarsening Factor": 2,
                             "Eliminate Tube Pressures With Velocities": True}}
A = -hb.galeri.create_matrix("Stokes-C", 3, nx)
P = hb.Preconditioner(A, params, hb.galeri.create_testvector(A))
P.SetRank(rank, nranks)
P.Initialize()
print("owned", len(P.OwnedSubdomains()), "of", P.NumMySubdomains(0), flush=True)
P.Compute()
x = P.ApplyInverse(np.ones(A.shape[0]))
print("done", np.isfinite(x).all(), flush=True)
