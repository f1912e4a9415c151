"""GPU parity of the bordered variant: Preconditioner::SetBorder / ComputeBorder /
bordered ApplyInverse (src/HYMLS_Preconditioner.cpp:519-588,844-918,930-1070), SchurPreconditioner's and
CoarseSolver's bordered parts, and the BorderedSolver Krylov path (src/HYMLS_BorderedSolver.cpp:159-219),
against the CPU oracle on the same inputs."""
import numpy as np
import pytest
import scipy.sparse as sp

import hymls_b200 as hb
from oracle import hymls as oh, krylov as ok
from tests.common import make_params
from tests.conftest import load_fixture
from tests.test_gpu_parity import dictify, rel, TOL_STOKES

pytestmark = pytest.mark.gpu


def const_pressure(n, dof):
    # MainUtils::create_nullspace "Constant P" (src/HYMLS_MainUtils.cpp:378-394, normalised :431-439)
    V = np.zeros((n, 1))
    V[dof - 1::dof, 0] = 1.0
    return V / np.linalg.norm(V)


def build_bordered(p, A, V, W=None, Cm=None, solver=None):
    A = sp.csr_matrix(A)
    tv = hb.galeri.create_testvector(A)
    pd = dictify(p)
    if solver:
        pd["Solver"] = solver
    P = hb.Preconditioner(A, pd, tv)
    P.Initialize()
    P.SetBorder(V, W, Cm)
    P.Compute()
    O = oh.Preconditioner(A, p.copy(), tv)
    O.initialize()
    O.set_border(V, W, Cm)
    O.compute()
    return P, O


def cases():
    out = []
    A2, _, _ = load_fixture("cavity2d_32_Re0")
    n2 = A2.shape[0]
    # integration_tests/bordering2.xml: Cartesian sx=4, 2 levels, constant-pressure border, no fixed pressure
    out.append(("bordering2", make_params("Stokes-C", 2, 32, 4, 2, Fix_Pressure_Level=False), A2,
                const_pressure(n2, 3), None, None))
    # one level, and the exact path (Number of Levels = 0: border of the dense Schur complement)
    out.append(("one-level", make_params("Stokes-C", 2, 32, 4, 1, Fix_Pressure_Level=False), A2,
                const_pressure(n2, 3), None, None))
    out.append(("exact", make_params("Stokes-C", 2, 32, 8, 0, Fix_Pressure_Level=False), A2,
                const_pressure(n2, 3), None, None))
    # three levels, two border columns, W != V and C != 0 on a non-singular matrix
    rng = np.random.default_rng(11)
    V = np.concatenate([const_pressure(n2, 3), rng.uniform(-1, 1, (n2, 1)) / np.sqrt(n2)], axis=1)
    W = np.concatenate([rng.uniform(-1, 1, (n2, 1)) / np.sqrt(n2), const_pressure(n2, 3)], axis=1)
    Cm = np.array([[0.5, -0.25], [0.125, 2.0]])
    out.append(("m2-general", make_params("Stokes-C", 2, 32, 4, 3, 2), A2, V, W, Cm))
    # 3D, skew partitioner (the reference's 3D Stokes configuration), constant-pressure border
    A3, _, _ = load_fixture("cavity3d_16_Re0")
    out.append(("3d-skew", make_params("Stokes-C", 3, 16, 4, 2, 2, Partitioner="Skew Cartesian",
                                       Fix_Pressure_Level=False), A3, const_pressure(A3.shape[0], 4), None, None))
    return out


@pytest.mark.parametrize("name,p,A,V,W,Cm", cases(), ids=[c[0] for c in cases()])
def test_bordered_apply_inverse_matches_oracle(name, p, A, V, W, Cm):
    P, O = build_bordered(p, A, V, W, Cm)
    n, m = V.shape
    rng = np.random.default_rng(2)
    B = rng.uniform(-1, 1, (n, 2))
    T = rng.uniform(-1, 1, (m, 2))
    X, S = P.ApplyInverseBordered(B, T)
    Xo, So = O.apply_inverse_bordered(B, T)
    scale = np.linalg.norm(np.concatenate([Xo, So]))
    assert np.linalg.norm(X - Xo) / scale < TOL_STOKES
    assert np.linalg.norm(S - So) / np.linalg.norm(So) < 1e-7    # S is O(1) while X is O(1e5): own scale
    # Epetra_Operator::ApplyInverse with a border set: T = 0, S discarded (Preconditioner.cpp:594-605)
    x = P.ApplyInverse(B[:, 0])
    assert rel(x, O.apply_inverse(B[:, 0])) < TOL_STOKES
    # linearity in (B, T)
    X2, S2 = P.ApplyInverseBordered(2.0 * B[:, :1] - 3.0 * B[:, 1:], 2.0 * T[:, :1] - 3.0 * T[:, 1:])
    assert np.linalg.norm(X2[:, 0] - (2.0 * X[:, 0] - 3.0 * X[:, 1])) / scale < 1e-11
    assert np.linalg.norm(S2[:, 0] - (2.0 * S[:, 0] - 3.0 * S[:, 1])) / scale < 1e-11


def test_bordered_gmres_reference_target():
    """integration_tests/bordering2.xml on the shipped 32x32/Re0 fixture: left-preconditioned bordered GMRES
    from zero, tol 1e-10; target <= 68 iterations, residual and error (constant pressure projected out)
    <= 5e-8; the GPU must agree with the oracle's count within +-1."""
    A, b, sol = load_fixture("cavity2d_32_Re0")
    n = A.shape[0]
    p = make_params("Stokes-C", 2, 32, 4, 2, Fix_Pressure_Level=False)
    V = const_pressure(n, 3)
    solver = {"Krylov Method": "GMRES", "Initial Vector": "Zero", "Left or Right Preconditioning": "Left",
              "Use Bordering": True,
              "Iterative Solver": {"Maximum Iterations": 100, "Maximum Restarts": 1, "Convergence Tolerance": 1e-10}}
    P, O = build_bordered(p, A, V, solver=solver)
    S = hb.Solver(P)
    x = S.ApplyInverse(b)

    def op(v):
        return np.concatenate([A @ v[:n] + V @ v[n:], V.T @ v[:n]])

    def pm(v):
        X, Sb = O.apply_inverse_bordered(v[:n], v[n:])
        return np.concatenate([X[:, 0], Sb[:, 0]])

    xo, its, conv, h = ok.gmres(op, np.concatenate([b, [0.0]]), np.zeros(n + 1), pm, side="Left", tol=1e-10,
                                max_iters=100, max_restarts=1)
    assert S.info["converged"] and conv
    assert S.num_iter <= 68 and abs(S.num_iter - its) <= 1
    k = min(len(h), len(S.history), 15)
    assert np.allclose(S.history[:k], h[:k], rtol=1e-5)
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) <= 5e-8
    err = x - sol
    err -= V[:, 0] * (V[:, 0] @ err)
    assert np.linalg.norm(err) / np.linalg.norm(b) <= 5e-8


def test_bordered_gmres_right_preconditioning_3d():
    A, b, sol = load_fixture("cavity3d_16_Re0")
    n = A.shape[0]
    p = make_params("Stokes-C", 3, 16, 4, 2, 2, Partitioner="Skew Cartesian", Fix_Pressure_Level=False)
    V = const_pressure(n, 4)
    solver = {"Krylov Method": "GMRES", "Initial Vector": "Zero", "Left or Right Preconditioning": "Right",
              "Iterative Solver": {"Maximum Iterations": 200, "Maximum Restarts": 1, "Convergence Tolerance": 1e-8}}
    P, O = build_bordered(p, A, V, solver=solver)
    S = hb.Solver(P)
    x = S.ApplyInverse(b)

    def op(v):
        return np.concatenate([A @ v[:n] + V @ v[n:], V.T @ v[:n]])

    def pm(v):
        X, Sb = O.apply_inverse_bordered(v[:n], v[n:])
        return np.concatenate([X[:, 0], Sb[:, 0]])

    xo, its, conv, h = ok.gmres(op, np.concatenate([b, [0.0]]), np.zeros(n + 1), pm, side="Right", tol=1e-8,
                                max_iters=200, max_restarts=1)
    assert S.info["converged"] and conv and abs(S.num_iter - its) <= 1
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) <= 1e-6
    assert abs(V[:, 0] @ x) <= 1e-6 * np.linalg.norm(x)     # the border row: V' x = 0


def test_remove_border_restores_plain_preconditioner():
    A = -hb.galeri.create_matrix("Stokes-C", 2, 16)
    p = make_params("Stokes-C", 2, 16, 4, 1)
    V = const_pressure(A.shape[0], 3)
    P, O = build_bordered(p, A, V)
    v = np.random.default_rng(4).uniform(-1, 1, A.shape[0])
    with_border = P.ApplyInverse(v)
    P.SetBorder(None)
    with pytest.raises(hb.HymlsError) as e:      # "Compute() needs to be called after SetBorder" (:874-875)
        P.ApplyInverse(v)
    assert e.value.code == -2
    P.Compute()
    O.set_border(None)
    O.compute()
    plain = P.ApplyInverse(v)
    assert rel(plain, O.apply_inverse(v)) < TOL_STOKES
    assert rel(with_border, plain) > 1e-6
