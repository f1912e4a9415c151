"""The reference's own run-time invariants (HYMLS::Tester, src/HYMLS_Tester.cpp, active under HYMLS_TESTING)
restated as tests on the generators and the shipped fixtures: isFmatrix (:204-250).  isDDcorrect (:253-455)
is tests/test_host_maps.py::test_domain_decomposition_decouples_interiors."""
import numpy as np
import scipy.sparse as sp

from oracle import galeri
from tests.common import make_params
from tests.conftest import load_fixture

FLOAT_TOL = 256 * np.finfo(float).eps   # Tester::float_tol(), src/HYMLS_Tester.hpp:92


def is_f_matrix(A, dof, pvar):
    """every velocity row couples to at most two pressures, with opposite values (a discrete gradient).
    (The reference also asks for a symmetric graph; the Dirichlet velocity rows of the generators and fixtures
    are stored diagonal-only, so that part is not asserted here.)"""
    A = sp.csr_matrix(A)
    cols_is_p = (np.arange(A.shape[1]) % dof) == pvar
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    sel = cols_is_p[A.indices] & ((rows % dof) != pvar) & (A.data != 0)
    cnt = np.bincount(rows[sel], minlength=A.shape[0])
    psum = np.bincount(rows[sel], weights=A.data[sel], minlength=A.shape[0])
    return bool((cnt <= 2).all() and (np.abs(psum) <= FLOAT_TOL * max(1.0, abs(A).max())).all())


def test_generators_and_fixtures_are_f_matrices():
    assert is_f_matrix(galeri.stokes(16, 16, 1, 2, 256.0, 1.0), 3, 2)
    assert is_f_matrix(galeri.stokes(8, 8, 8, 3, 64.0, 1.0), 4, 3)
    for name, dof in (("cavity2d_32_Re0", 3), ("cavity2d_32_Re1000", 3), ("cavity3d_16_Re0", 4)):
        A, _, _ = load_fixture(name)
        assert is_f_matrix(A, dof, dof - 1), name
    # a Laplace matrix with the same dof layout is not one (the check has teeth)
    L = sp.kron(galeri.create_matrix(make_params("Laplace", 2, 8, 4, 1).sublist("Problem")), np.ones((3, 3)))
    assert not is_f_matrix(L, 3, 2)
