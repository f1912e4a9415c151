"""The BASELINE.json configurations at their STATED sizes: the CUDA path against the C++/OpenMP restatement of the
reference (oracle/cpp, itself checked against the numpy oracle in tests/test_oracle_cpp.py) on the same inputs --
ApplyInverse within the north star's 1e-12, Krylov iteration count within +-1 and the first 15 entries of the
residual history.  The oracle runs with one step of extended-precision iterative refinement on every direct solve
(ho_set_refinement): its plain FP64 sparse-LU solves are themselves 1e-12 (16^3) .. 1e-10 (64^3) away from the
exact-arithmetic preconditioner, the refined ones 1e-15 (checked against oracle/extended.py in
tests/test_oracle_cpp.py), so 1e-12 is a meaningful bound at every size.

  config 1  testSuite/laplace.xml    Laplace2D 128x128, sx=4, 2 levels, CG, tol 1e-10
  config 2  testSuite/stokes2D.xml   Stokes2D 128x128, Skew Cartesian, sx=4, 2 levels, GMRES(30), tol 1e-10
  config 4  testSuite/cavity3D.xml   scaled to 64^3 (synthetic Stokes3D), sx=4, 2 levels, GMRES, Skew Cartesian
                                      (the reference's 3D Stokes partitioner)
Config 3 (cavity.xml, bordered, shipped 64x64 Re1000 fixture) is covered by tests/test_gpu_bordered.py and
tests/test_driver.py against the numpy oracle (the C++ oracle has no bordering); config 5 (128^3) is compared in
profiles/r02_parity_128cube.md (a 4-minute CPU run, not a test).  Both sides start from the same initial vector."""
import numpy as np
import pytest
import scipy.sparse as sp

import hymls_b200 as hb
from oracle import cpp_oracle as oc
from oracle.params import ParameterList
from tests.test_gpu_parity import TRUTH_TOL

pytestmark = pytest.mark.gpu


def _pl(d):
    pl = ParameterList()
    for k, v in d.items():
        pl[k] = _pl(v) if isinstance(v, dict) else v
    return pl


CONFIGS = [
    ("laplace.xml", "Laplace", 2, 128, {"Separator Length": 4, "Number of Levels": 2},
     {"Krylov Method": "CG", "tol": 1e-10, "blocks": 300, "restarts": 20}, TRUTH_TOL),
    ("stokes2D.xml", "Stokes-C", 2, 128, {"Partitioner": "Skew Cartesian", "Separator Length": 4, "Number of Levels": 2},
     {"Krylov Method": "GMRES", "tol": 1e-10, "blocks": 30, "restarts": 20}, TRUTH_TOL),
    ("cavity3D.xml @ 64^3", "Stokes-C", 3, 64,
     {"Partitioner": "Skew Cartesian", "Separator Length": 4, "Number of Levels": 2},
     {"Krylov Method": "GMRES", "tol": 1e-8, "blocks": 300, "restarts": 20}, TRUTH_TOL),
]


@pytest.mark.parametrize("name,eqn,dim,nx,prec,sol,tol", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_baseline_config_at_stated_size(name, eqn, dim, nx, prec, sol, tol):
    A = hb.galeri.create_matrix(eqn, dim, nx)
    if eqn == "Stokes-C":
        A = -A
    A = sp.csr_matrix(A)
    tv = hb.galeri.create_testvector(A)
    n = A.shape[0]
    params = {"Problem": {"Equations": eqn, "Dimension": dim, "nx": nx, "ny": nx, "nz": nx if dim == 3 else 1},
              "Preconditioner": dict(prec),
              "Solver": {"Krylov Method": sol["Krylov Method"], "Initial Vector": "Previous",
                         "Left or Right Preconditioning": "Right",
                         "Iterative Solver": {"Maximum Iterations": 500, "Convergence Tolerance": sol["tol"],
                                              "Num Blocks": sol["blocks"], "Maximum Restarts": sol["restarts"],
                                              "Implicit Residual Scaling": "Norm of Initial Residual"}}}
    P = hb.Preconditioner(A, params, tv)
    P.Initialize()
    P.Compute()
    O = oc.Preconditioner(A, _pl(params), tv, oc.maps_from_library(P), refine_steps=1)
    O.compute()
    rng = np.random.default_rng(42)
    xex = rng.uniform(-1, 1, n)
    b = A @ xex
    x0 = np.random.default_rng(43).uniform(-1, 1, n)
    # ApplyInverse at the stated size
    xg, xo = P.ApplyInverse(b), O.apply_inverse(b)
    assert np.linalg.norm(xg - xo) <= tol * np.linalg.norm(xo)
    # Krylov solve: iteration count and residual history
    S = hb.Solver(P)
    xs = S.ApplyInverse(b, x=x0.copy())
    xc, its, conv, hist, _ = O.solve(b, x0=x0, method=sol["Krylov Method"], tol=sol["tol"], max_iters=500,
                                     num_blocks=sol["blocks"], max_restarts=sol["restarts"])
    assert S.info["converged"] and conv
    assert abs(S.num_iter - its) <= 1, (S.num_iter, its)
    k = min(15, len(hist), len(S.history))
    assert np.allclose(S.history[:k], hist[:k], rtol=1e-6), (S.history[:k], hist[:k])
    assert np.linalg.norm(A @ xs - b) <= 10 * sol["tol"] * np.linalg.norm(b)
