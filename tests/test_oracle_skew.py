"""Pins the skew partitioner oracle to the reference's unit tests
(testSuite/unit_tests/HYMLS_SkewCartesianPartitioner.cpp) and integration targets (stokes*_3D.xml)."""
import numpy as np
import pytest

from oracle.params import ParameterList
from oracle.skew import SkewCartesianPartitioner


def _params(nx, ny, nz, sx, eqn=None, dof=None, dim=None, perio=False):
    p = ParameterList()
    pr = p.sublist("Problem")
    pr.set("nx", nx); pr.set("ny", ny); pr.set("nz", nz)
    if dim is not None:
        pr.set("Dimension", dim)
    if dof is not None:
        pr.set("Degrees of Freedom", dof)
    if eqn is not None:
        pr.set("Equations", eqn)
    if perio:
        pr.set("x-periodic", True); pr.set("y-periodic", True); pr.set("z-periodic", True)
    p.sublist("Preconditioner").set("Separator Length", sx)
    return p


def test_operator_ids():  # :43-69 (4 fake ranks, 8^3, dof 3, sx 4)
    part = SkewCartesianPartitioner(_params(8, 8, 8, 4, dof=3), 0, 4, 0).partition()
    exp = {(0, 0, 0): 0, (0, 1, 0): 2, (7, 0, 0): 4, (3, 4, 0): 8, (3, 4, 3): 20, (3, 4, 4): 20, (0, 0, 4): 12,
           (7, 7, 7): 21}
    for (i, j, k), sd in exp.items():
        assert part(i, j, k) == sd


def test_pid():  # :71-95
    part = SkewCartesianPartitioner(_params(8, 8, 8, 4, dof=3), 0, 4, 0).partition()

    def pid(i, j, k):
        return part.pid_map[part(i, j, k)]
    assert (pid(0, 0, 0), pid(0, 1, 0), pid(7, 0, 0), pid(7, 7, 7)) == (0, 1, 1, 3)


def test_get_subdomain_roundtrip_periodic():  # :97-141
    nx, cl = 12, 6
    part = SkewCartesianPartitioner(_params(nx, nx, nx, cl, eqn="Stokes-C", perio=True), 0, 1, 0).partition()
    for sd in range(part.num_local_parts()):
        gsd = part.sd_map[sd]
        i, j, k = part.subdomain_position(gsd, cl, cl, cl)
        i, j, k = (i % nx + nx) % nx, (j % nx + nx) % nx, (k % nx + nx) % nx
        assert part.subdomain_id(cl, cl, cl, i, j, k) == gsd


def test_subdomain0_at_origin():  # :143-171
    nx = 4
    while nx <= 32:
        cl = 2
        while cl <= nx:
            part = SkewCartesianPartitioner(_params(nx, nx, nx, cl, eqn="Stokes-C"), 0, 4, 0)
            assert part.subdomain_position(0, cl, cl, cl) == (0, 0, 0)
            cl *= 2
        nx *= 2


def test_num_global_parts():  # :173-196
    part = SkewCartesianPartitioner(_params(32, 28, 24, 4, eqn="Stokes-C"), 0, 1, 0).partition()
    assert part.num_global_parts(4, 4, 4) == part.num_local_parts()


def _coverage(part):
    n = part.nx * part.ny * part.nz * part.dof
    seen = np.zeros(n, dtype=bool)
    for sd in range(part.num_local_parts()):
        it, gr = part.get_groups(sd)
        seen[it] = True
        for _, nodes in gr:
            seen[nodes] = True
    return seen


def test_2d_nodes_covered():  # :198-236
    part = SkewCartesianPartitioner(_params(8, 8, 1, 4, eqn="Stokes-C", dim=2), 0, 1, 0).partition()
    assert _coverage(part).all()


def test_one_pressure_separator_per_domain_2d():  # :238-272
    part = SkewCartesianPartitioner(_params(8, 8, 1, 4, eqn="Stokes-C", dim=2), 0, 1, 0).partition()
    for sd in range(part.num_local_parts()):
        _, gr = part.get_groups(sd)
        assert sum(1 for _, nodes in gr for g in nodes if g % 3 == 2) == 1


def test_3d_nodes_covered():  # :274-311
    part = SkewCartesianPartitioner(_params(8, 8, 8, 4, eqn="Stokes-C", dof=4), 0, 1, 0).partition()
    assert _coverage(part).all()


def test_5dof_nodes_covered():  # :313-
    part = SkewCartesianPartitioner(_params(8, 8, 8, 4, eqn="Bous-C"), 0, 1, 0).partition()
    assert _coverage(part).all()


# ---- integration targets of the reference that use the skew partitioner (fixtures under tests/golden) ----
from oracle import galeri, hymls, krylov  # noqa: E402
from tests.common import make_params  # noqa: E402
from tests.conftest import load_fixture  # noqa: E402


def _solve(name, nx, dim, p, tol, max_iters, explicit=True):
    A, b, sol = load_fixture(name)
    n = A.shape[0]
    dof = dim + 1
    prec = hymls.Preconditioner(A, p, galeri.create_testvector(A))
    prec.initialize()
    prec.compute()
    x0 = np.random.default_rng(43).uniform(-1, 1, n)
    x, its, conv, _ = krylov.gmres(lambda v: A @ v, b, x0, prec.apply_inverse, side="Right", tol=tol,
                                   max_iters=max_iters, max_restarts=1, explicit_test=explicit,
                                   imp_scaling="Norm of RHS", exp_scaling="Norm of RHS")
    pv = np.zeros(n); pv[dof - 1::dof] = 1; pv /= np.linalg.norm(pv)
    err = x - sol
    err -= pv * (pv @ err)
    return its, conv, np.linalg.norm(A @ x - b) / np.linalg.norm(b), np.linalg.norm(err) / np.linalg.norm(b)


def test_stokes0_3d_exact():  # integration_tests/stokes0_3D.xml: Skew sx=8, 0 levels -> 1 iteration
    p = make_params("Stokes-C", 3, 16, 8, 0, Partitioner="Skew Cartesian")
    its, conv, res, err = _solve("cavity3d_16_Re0", 16, 3, p, 1e-8, 5)
    assert conv and its == 1 and res <= 1e-8 and err <= 1e-8


def test_stokes1_3d_target():  # stokes1_3D.xml: Skew sx=8, 1 level, tol 1e-8 -> <= 130 its, <= 1.5e-5
    p = make_params("Stokes-C", 3, 16, 8, 1, Partitioner="Skew Cartesian")
    its, conv, res, err = _solve("cavity3d_16_Re0", 16, 3, p, 1e-8, 160)
    assert conv and its <= 130 and res <= 1.5e-5 and err <= 1.5e-5


def test_stokes2_3d_target():  # stokes2_3D.xml: Skew sx=4, cx=2, 2 levels, velocities not linked -> <= 145 its
    p = make_params("Stokes-C", 3, 16, 4, 2, 2, Partitioner="Skew Cartesian", Eliminate_Velocities_Together=False)
    its, conv, res, err = _solve("cavity3d_16_Re0", 16, 3, p, 1e-8, 160)
    assert conv and its <= 145 and res <= 1e-5 and err <= 1e-5


def test_stokes1_2d_target():  # stokes1.xml: Skew sx=4, 1 level, right-preconditioned from zero, tol 1e-6 -> <= 23 its
    A, b, sol = load_fixture("cavity2d_32_Re0")
    p = make_params("Stokes-C", 2, 32, 4, 1, Partitioner="Skew Cartesian")
    prec = hymls.Preconditioner(A, p, galeri.create_testvector(A))
    prec.initialize(); prec.compute()
    x, its, conv, _ = krylov.gmres(lambda v: A @ v, b, np.zeros(A.shape[0]), prec.apply_inverse, side="Right",
                                   tol=1e-6, max_iters=100, max_restarts=1)
    assert conv and its <= 23 and np.linalg.norm(A @ x - b) / np.linalg.norm(b) <= 5e-6


@pytest.mark.parametrize("dim,sx", [(2, 4), (2, 6), (2, 8), (3, 4), (3, 6), (3, 8), (3, 10)])
def test_template_is_a_polytope_in_sheared_coordinates(dim, sx):
    """hymls_b200/csrc/partitioner.cpp (SkewShape) describes the subdomain template geometrically: per variable
    type the lattice points of a polytope bounded by planes of constant s = x+y, d = x-y, t = x-y+z and z.  The
    restated reference construction (oracle/skew.py::_template, src/HYMLS_SkewCartesianPartitioner.cpp:372-565)
    must produce exactly those node sets."""
    from oracle.skew import SkewCartesianPartitioner
    from tests.common import stokes_var_params
    n = 4 * sx
    part = SkewCartesianPartitioner(stokes_var_params(dim, n, n, n if dim == 3 else 1, sx), 0, 1, 0).partition()
    dof, w = part.dof, 4 * sx
    got = {v: set() for v in range(dof)}
    for cat in part.groups_template:
        for grp in cat:
            for node in grp:
                got[node % dof].add(((node // dof) % w - 1 - sx, (node // dof // w) % w - 1 - 3 * sx // 2,
                                     node // dof // w // w - 2 * sx))
    b = sx

    def inside(kind, x, y, z):
        if dim == 2 and z != 0:
            return False
        s, d, t = x + y, x - y, x - y + z
        if kind == "P":
            return -1 <= s <= b - 2 and 0 <= d <= b - 1 and 0 <= t <= b - 1 and abs(z) <= b - 1
        if kind == "U":
            return -2 <= s <= b - 2 and -1 <= d <= b - 1 and -1 <= t <= b - 1 and abs(z) <= b - 1
        if kind == "V":
            return -2 <= s <= b - 2 and 0 <= d <= b and 0 <= t <= b and abs(z) <= b - 1
        ok = -2 <= s <= b - 1 and -1 <= d <= b and -1 <= t <= b - 1 and -b <= z <= b - 1
        return ok and (t % 2 == 1 or not (s in (-2, b - 1) or d in (-1, b)))
    kinds = ["U", "V", "W", "P"] if dim == 3 else ["U", "V", "P"]
    r = range(-2 * sx, 2 * sx + 1)
    for v, kind in enumerate(kinds):
        want = {(x, y, z) for z in (r if dim == 3 else [0]) for y in r for x in r if inside(kind, x, y, z)}
        assert got[v] == want, (kind, len(got[v]), len(want))
