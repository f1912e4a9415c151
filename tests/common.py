"""Shared helpers for tests: parameter lists in the shape of the reference's XML files."""
from oracle.params import ParameterList


def make_params(eqn, dim, nx, sx, levels, cx=None, ny=None, nz=None, **prec):
    p = ParameterList()
    pr = p.sublist("Problem")
    pr.set("Equations", eqn)
    pr.set("Dimension", dim)
    pr.set("nx", nx)
    pr.set("ny", nx if ny is None else ny)
    pr.set("nz", (nx if dim == 3 else 1) if nz is None else nz)
    pc = p.sublist("Preconditioner")
    pc.set("Partitioner", "Cartesian")
    pc.set("Separator Length", sx)
    pc.set("Number of Levels", levels)
    if cx:
        pc.set("Coarsening Factor", cx)
    for k, v in prec.items():
        pc.set(k.replace("_", " "), v)
    return p


def stokes_var_params(dim, nx, ny, nz, sx, cf=2):
    """Parameter list of unit_tests/HYMLS_OverlappingPartitioner.cpp:355-376 / :551-573."""
    p = ParameterList()
    pr = p.sublist("Problem")
    pr.set("nx", nx)
    pr.set("ny", ny)
    pr.set("nz", nz)
    for i in range(dim):
        pr.sublist("Variable %d" % i).set("Variable Type", "Velocity")
    pr.sublist("Variable %d" % dim).set("Variable Type", "Pressure")
    pr.set("Dimension", dim)
    pr.set("Degrees of Freedom", dim + 1)
    pc = p.sublist("Preconditioner")
    pc.set("Separator Length", sx)
    pc.set("Coarsening Factor", cf)
    return p
