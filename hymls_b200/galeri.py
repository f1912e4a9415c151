"""Synthetic structured-grid Jacobians (vectorised numpy), the inputs of the headline benchmarks.

Same matrices as the reference's generators GaleriExt::Matrices::Stokes2D/Stokes3D (C grid,
src/GaleriExt_Stokes2D.h:87-218, src/GaleriExt_Stokes3D.h:89-285), Darcy2D/3D and Galeri Cross2D/3D, as
called by MainUtils::create_matrix (src/HYMLS_MainUtils.cpp:260-348), including the explicit zeros the
reference stores for couplings to boundary velocities.  Built array-at-a-time so that 128^3 (8.4 M rows)
takes seconds; tests compare it entry by entry with the row-by-row oracle and the shipped fixtures.
"""
import numpy as np
import scipy.sparse as sp


def _grid(nx, ny, nz):
    c = np.arange(nx * ny * nz, dtype=np.int64)
    return c, c % nx, (c // nx) % ny, c // (nx * ny)


def laplace(nx, ny, nz=1, dim=2):
    """Galeri Laplace2D / Laplace3D scaled by -1 (create_matrix :341-346)."""
    c, i, j, k = _grid(nx, ny, nz)
    rows, cols, vals = [c], [c], [np.full(len(c), -(4.0 if dim == 2 else 6.0))]
    for ok, off in ((i > 0, -1), (i < nx - 1, 1), (j > 0, -nx), (j < ny - 1, nx)) + \
            (((k > 0, -nx * ny), (k < nz - 1, nx * ny)) if dim == 3 else ()):
        rows.append(c[ok]); cols.append(c[ok] + off); vals.append(np.ones(int(ok.sum())))
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(len(c), len(c)))
    A.sort_indices()
    return A


def stokes(nx, ny, nz, dim, a, b):
    """GaleriExt::Stokes2D / Stokes3D (non-periodic, C grid) with A-part scale a and B-part scale b."""
    dof = dim + 1
    pv = dim
    c, ci, cj, ck = _grid(nx, ny, nz)
    n = len(c) * dof
    idx = (ci, cj, ck)
    ext = (nx, ny, nz)
    stride = (1, nx, nx * ny)
    R, Cc, V = [], [], []

    def add(r, col, v):
        R.append(r); Cc.append(col); V.append(np.broadcast_to(v, r.shape).astype(np.float64))

    # pressure rows (Darcy with a=0, b -> -b, so c = b): -c u_self (if fwd exists), +c u_back
    for d in range(dim):
        fwd = idx[d] < ext[d] - 1
        bwd = idx[d] > 0
        add(c[fwd] * dof + pv, c[fwd] * dof + d, -b)
        add(c[bwd] * dof + pv, (c[bwd] - stride[d]) * dof + d, b)
    # velocity rows
    for d in range(dim):
        row = c * dof + d
        fwd = idx[d] < ext[d] - 1            # forward neighbour exists
        fwd2 = idx[d] < ext[d] - 2           # ... and its forward neighbour too
        # gradient part: +b p_self, -b p_fwd
        add(row[fwd], c[fwd] * dof + pv, b)
        add(row[fwd], (c[fwd] + stride[d]) * dof + pv, -b)
        # Dirichlet rows on the far boundary: diagonal only
        diag = np.empty(len(c))
        if dim == 2:
            diag[~fwd] = -(b / (a * a)) * a
        else:
            diag[~fwd] = -(-1.0 / a) * a
        # interior rows: -(lap*a + add_to_diag)
        lapdiag = 4.0 if dim == 2 else 6.0
        add_to_diag = np.zeros(len(c))
        trans = [t for t in range(dim) if t != d]
        for t in trans:
            on_bnd = (idx[t] == 0) | (idx[t] == ext[t] - 1)
            if dim == 2:
                add_to_diag = np.where(on_bnd, a, add_to_diag)
            else:
                add_to_diag = add_to_diag + np.where(on_bnd, a, 0.0)
        diag[fwd] = -(lapdiag * a + add_to_diag[fwd])
        add(row, row, diag)
        # off-diagonal Laplace couplings (+a), only for rows that are not Dirichlet rows
        for t in range(dim):
            for sgn in (-1, 1):
                ok = fwd & ((idx[t] > 0) if sgn < 0 else (idx[t] < ext[t] - 1))
                val = np.full(int(ok.sum()), a)
                if t == d and sgn > 0:
                    # coupling to the velocity on the boundary is stored as an explicit zero
                    val = np.where(fwd2[ok], a, 0.0)
                add(row[ok], (c[ok] + sgn * stride[t]) * dof + d, val)
    A = sp.csr_matrix((np.concatenate(V), (np.concatenate(R), np.concatenate(Cc))), shape=(n, n))
    A.sort_indices()
    return A


def create_matrix(eqn, dim, nx, ny=None, nz=None):
    """MainUtils::create_matrix for 'Laplace' and 'Stokes-C' (a = nx^2, b = 1)."""
    ny = nx if ny is None else ny
    nz = (nx if dim == 3 else 1) if nz is None else nz
    if eqn == "Laplace":
        return laplace(nx, ny, nz, dim)
    if eqn == "Stokes-C":
        return stokes(nx, ny, nz, dim, float(nx * nx), 1.0)
    raise ValueError(eqn)


def create_testvector(A):
    """MainUtils::create_testvector (:208-258): ones, zero on rows whose only non-zero values are diagonal."""
    A = sp.csr_matrix(A)
    n = A.shape[0]
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    off = (A.data != 0) & (A.indices != rows)
    has_off = np.zeros(n, dtype=bool)
    has_off[rows[off]] = True
    tv = np.ones(n)
    tv[~has_off] = 0.0
    return tv
