"""Matrix-generator oracle (ORACLE -- test infrastructure only).

Literal row-by-row restatement of
  * Galeri Cross2D/Cross3D (third-party Trilinos/Galeri; 5/7-point stencils with
    Dirichlet truncation; call sites src/GaleriExt_Stokes2D.h:82, GaleriExt_Stokes3D.h:82)
  * GaleriExt::GetNeighboursCartesian{2d,3d}    src/GaleriExt_Periodic.cpp:9-60
  * GaleriExt::Matrices::Darcy2D / Darcy3D      src/GaleriExt_Darcy2D.h, GaleriExt_Darcy3D.h:46-178
  * GaleriExt::Matrices::Stokes2D / Stokes3D    src/GaleriExt_Stokes2D.h:87-218, GaleriExt_Stokes3D.h:89-285
  * MainUtils::create_matrix / create_testvector src/HYMLS_MainUtils.cpp:208-348
Explicit zeros written by the reference (couplings to boundary velocities) are kept.
"""
import numpy as np
import scipy.sparse as sp

X_PERIO, Y_PERIO, Z_PERIO = 1, 2, 4


def neighbours2d(i, nx, ny, perio=0):
    ix = i % nx
    iy = (i - ix) // nx
    left = -1 if ix == 0 else i - 1
    right = -1 if ix == nx - 1 else i + 1
    lower = -1 if iy == 0 else i - nx
    upper = -1 if iy == ny - 1 else i + nx
    if perio & X_PERIO:
        left = iy * nx + (ix - 1) % nx
        right = iy * nx + (ix + 1) % nx
    if perio & Y_PERIO:
        lower = ((iy - 1) % ny) * nx + ix
        upper = ((iy + 1) % ny) * nx + ix
    return left, right, lower, upper


def neighbours3d(i, nx, ny, nz, perio=0):
    ixy = i % (nx * ny)
    iz = (i - ixy) // (nx * ny)
    if not (perio & Z_PERIO):
        below = -1 if iz == 0 else i - nx * ny
        above = -1 if iz == nz - 1 else i + nx * ny
    else:
        below = (i - nx * ny) % (nx * ny * nz)
        above = (i + nx * ny) % (nx * ny * nz)
    left, right, lower, upper = neighbours2d(ixy, nx, ny, perio)
    off = iz * nx * ny
    left, right, lower, upper = [(-1 if v == -1 else v + off) for v in (left, right, lower, upper)]
    return left, right, lower, upper, below, above


class _Rows:
    def __init__(self, n):
        self.n = n
        self.r, self.c, self.v = [], [], []

    def insert(self, row, cols, vals):
        self.r.extend([row] * len(cols))
        self.c.extend(cols)
        self.v.extend(vals)

    def csr(self):
        # keep explicit zeros: build CSR by hand (coo->csr would also keep them, but sums dups)
        m = sp.coo_matrix((np.asarray(self.v, dtype=np.float64),
                           (np.asarray(self.r, dtype=np.int64), np.asarray(self.c, dtype=np.int64))),
                          shape=(self.n, self.n)).tocsr()
        m.sort_indices()
        return m


def laplace_rows(nx, ny, nz, dim):
    """Cross2D(4,-1..)/Cross3D(6,-1..): list of (cols, vals) per cell, diagonal first."""
    rows = []
    for i in range(nx * ny * nz):
        if dim == 2:
            nb = neighbours2d(i, nx, ny)
            diag = 4.0
        else:
            nb = neighbours3d(i, nx, ny, nz)
            diag = 6.0
        # Galeri order: left, right, lower, upper, (below, above), then diagonal last
        cols = [j for j in nb if j != -1]
        vals = [-1.0] * len(cols)
        cols.append(i)
        vals.append(diag)
        rows.append((cols, vals))
    return rows


def laplace(nx, ny, nz=1, dim=2):
    """Galeri 'Laplace2D'/'Laplace3D' scaled by -1 (src/HYMLS_MainUtils.cpp:341-346)."""
    R = _Rows(nx * ny * nz)
    for i, (cols, vals) in enumerate(laplace_rows(nx, ny, nz, dim)):
        R.insert(i, cols, [-v for v in vals])
    return R.csr()


def darcy(nx, ny, nz, dim, a, b, perio=0):
    dof = dim + 1
    n = nx * ny * nz * dof
    c = -b
    R = _Rows(n)
    for gid in range(n):
        ibase = gid // dof
        ivar = gid - ibase * dof
        if dim == 2:
            left, right, lower, upper = neighbours2d(ibase, nx, ny, perio)
            below = above = -1
        else:
            left, right, lower, upper, below, above = neighbours3d(ibase, nx, ny, nz, perio)
        cols, vals = [], []
        pv = dof - 1
        if ivar != pv:
            cols.append(gid)
            vals.append(a)
            nb = (right, upper, above)[ivar]
            if nb != -1:
                cols += [ibase * dof + pv, nb * dof + pv]
                vals += [-b, b]
        else:
            for k, nb in enumerate((right, upper, above)[:dim]):
                if nb != -1:
                    cols.append(ibase * dof + k)
                    vals.append(-c)
            for k, nb in enumerate((left, lower, below)[:dim]):
                if nb != -1:
                    cols.append(nb * dof + k)
                    vals.append(c)
        R.insert(gid, cols, vals)
    return R.csr()


def stokes(nx, ny, nz, dim, a, b, perio=0):
    """GaleriExt::Stokes2D/Stokes3D on a C grid."""
    dof = dim + 1
    n = nx * ny * nz * dof
    D = darcy(nx, ny, nz, dim, 0.0, -b, perio)
    lap = laplace_rows(nx, ny, nz, dim)
    R = _Rows(n)
    for row in range(n):
        cols = list(D.indices[D.indptr[row]:D.indptr[row + 1]])
        vals = list(D.data[D.indptr[row]:D.indptr[row + 1]])
        len_darcy = len(cols)
        ivar = row % dof
        if ivar != dof - 1:
            row0 = row // dof
            lcols, lvals = list(lap[row0][0]), list(lap[row0][1])
            add = 0.0
            if dim == 2:
                left, right, lower, upper = neighbours2d(row0, nx, ny, perio)
                below = above = 0  # never -1 in 2D
            else:
                left, right, lower, upper, below, above = neighbours3d(row0, nx, ny, nz, perio)
            fwd = (right, upper, above)[ivar]
            fwd2 = -1
            if fwd > 0:
                if dim == 2:
                    fwd2 = neighbours2d(fwd, nx, ny, perio)[(1, 3)[ivar]]
                else:
                    fwd2 = neighbours3d(fwd, nx, ny, nz, perio)[(1, 3, 5)[ivar]]
            # pairs of transverse neighbours for the "centered" directions
            trans = {0: [(lower, upper), (below, above)],
                     1: [(left, right), (below, above)],
                     2: [(left, right), (lower, upper)]}[ivar]
            if dim == 2:
                trans = trans[:1]
            if fwd == -1:
                lcols = [row0]
                lvals = [b / (a * a)] if dim == 2 else [-1.0 / a]
            else:
                if dim == 2:
                    if trans[0][0] == -1 or trans[0][1] == -1:
                        add = a
                else:
                    for lo, hi in trans:
                        if lo == -1 or hi == -1:
                            add += a
            if fwd > 0 and fwd2 == -1:
                for j in range(len(lcols)):
                    if lcols[j] == fwd:
                        lvals[j] = 0.0
            for j in range(len(lcols)):
                c = lcols[j] * dof + ivar
                if c == row:
                    for k in range(len_darcy):
                        if cols[k] == c:
                            vals[k] = -(lvals[j] * a + add)
                else:
                    cols.append(c)
                    vals.append(-lvals[j] * a)
        R.insert(row, cols, vals)
    return R.csr()


def create_matrix(prob):
    """MainUtils::create_matrix (src/HYMLS_MainUtils.cpp:260-348) for Laplace / Stokes-C."""
    eqn = prob.get("Equations", "Laplace")
    dim = prob.get("Dimension", 2)
    nx = prob.get("nx", 32)
    ny = prob.get("ny", nx)
    nz = prob.get("nz", nx if dim > 2 else 1)
    if eqn == "Laplace":
        return laplace(nx, ny, nz, dim)
    if eqn == "Stokes-C":
        return stokes(nx, ny, nz, dim, float(nx * nx), 1.0)
    raise NotImplementedError(eqn)


def create_testvector(A):
    """MainUtils::create_testvector (src/HYMLS_MainUtils.cpp:208-258): ones, rows whose only
    non-zero VALUES are on the diagonal get 0."""
    A = sp.csr_matrix(A)
    n = A.shape[0]
    tv = np.ones(n)
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    off = (A.data != 0) & (A.indices != rows)
    has_off = np.zeros(n, dtype=bool)
    has_off[rows[off]] = True
    tv[~has_off] = 0.0
    return tv


def read_mtx(path):
    """MatrixMarket coordinate/array reader (keeps explicit zeros)."""
    with open(path) as f:
        header = f.readline().split()
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        if header[2] == "coordinate":
            m, n, nnz = [int(t) for t in line.split()]
            data = np.loadtxt(f, ndmin=2)
            r = data[:, 0].astype(np.int64) - 1
            c = data[:, 1].astype(np.int64) - 1
            v = data[:, 2]
            M = sp.coo_matrix((v, (r, c)), shape=(m, n)).tocsr()
            M.sort_indices()
            return M
        m, n = [int(t) for t in line.split()]
        data = np.loadtxt(f)
        return data.reshape((n, m)).T if n > 1 else data.reshape(m)
