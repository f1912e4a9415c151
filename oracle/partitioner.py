"""Index-set oracle (ORACLE -- test infrastructure only; never imported by the product).

Restates, in plain Python/numpy:
  * BasePartitioner::SetParameters        src/HYMLS_BasePartitioner.cpp:31-319
  * BasePartitioner::CreatePIDMap         src/HYMLS_BasePartitioner.cpp:361-586
  * CartesianPartitioner::GetGroups       src/HYMLS_CartesianPartitioner.cpp:224-408
  * HierarchicalMap::FillComplete / LinkSeparators / Spawn*
                                          src/HYMLS_HierarchicalMap.cpp:120-285,436-542
  * OverlappingPartitioner ctor / DetectSeparators / SpawnNextLevel
                                          src/HYMLS_OverlappingPartitioner.cpp:31-159
"""
import numpy as np

from .params import ParameterList

V_U, V_V, V_W, PRESSURE, INTERIOR = 0, 1, 2, 3, 4
X_PERIO, Y_PERIO, Z_PERIO = 1, 2, 4


def find_coarsening_factor(cx):
    # src/HYMLS_BasePartitioner.cpp:348-359
    b = 1
    while b < cx:
        for p in range(cx):
            if b ** p == cx:
                return b
        b += 1
    return cx


class CartesianPartitioner:
    """BasePartitioner + CartesianPartitioner for `nprocs` (fake) ranks, seen from `mypid`."""

    def __init__(self, params, level=0, nprocs=1, mypid=0):
        self.level = level
        self.nprocs_comm = nprocs
        self.mypid = mypid
        self.set_parameters(params)

    # -- src/HYMLS_BasePartitioner.cpp:31-319 ---------------------------------
    def set_parameters(self, params):
        prob = params.sublist("Problem")
        prec = params.sublist("Preconditioner")
        self.dim = prob.get("Dimension", 3)
        pvar = -1
        self.nx = prob.get("nx", -1)
        self.ny = prob.get("ny", self.nx)
        self.nz = prob.get("nz", self.nx if self.dim > 2 else 1)
        if self.nx == -1:
            raise ValueError("You must presently specify nx, ny (and possibly nz) in the 'Problem' sublist")
        xp = prob.get("x-periodic", False)
        yp = prob.get("y-periodic", False) if self.dim > 1 else False
        zp = prob.get("z-periodic", False) if self.dim > 2 else False
        perio = (X_PERIO if xp else 0) | (Y_PERIO if yp else 0) | (Z_PERIO if zp else 0)
        self.perio = prob.get("Periodicity", perio)

        sx = sy = -1
        sz = -1 if self.nz > 1 else 1
        if prec.isParameter("Separator Length (x)"):
            sx = prec.get("Separator Length (x)", sx)
        if prec.isParameter("Separator Length (y)"):
            sy = prec.get("Separator Length (y)", sy)
        if prec.isParameter("Separator Length (z)"):
            sz = prec.get("Separator Length (z)", sz)
        if sx == -1:
            sx = prec.get("Separator Length", 4)
        if sy == -1:
            sy = prec.get("Separator Length", sx)
        if sz == -1:
            sz = prec.get("Separator Length", sx)
        if sx <= 1:
            raise ValueError("Separator Length not set correctly")
        self.sx, self.sy, self.sz = sx, sy, sz

        cx = cy = -1
        cz = -1 if self.nz > 1 else 1
        if prec.isParameter("Coarsening Factor (x)"):
            cx = prec.get("Coarsening Factor (x)", cx)
        if prec.isParameter("Coarsening Factor (y)"):
            cy = prec.get("Coarsening Factor (y)", cy)
        if prec.isParameter("Coarsening Factor (z)"):
            cz = prec.get("Coarsening Factor (z)", cz)
        if cx == -1:
            cx = prec.get("Coarsening Factor", sx)
        if cy == -1:
            cy = prec.get("Coarsening Factor", cx)
        if cz == -1:
            cz = prec.get("Coarsening Factor", cx)
        if cx <= 1:
            raise ValueError("Coarsening Factor not set correctly")
        self.cx, self.cy, self.cz = cx, cy, cz

        r = {"x": -1, "y": -1, "z": -1}
        at_level = "Retain Nodes at Level %d" % self.level
        for d in "xyz":
            if prec.isParameter("Retain Nodes (%s)" % d):
                r[d] = prec.get("Retain Nodes (%s)" % d, r[d])
            if prec.isParameter(at_level + " (%s)" % d):
                r[d] = prec.get(at_level + " (%s)" % d, r[d])
        for d in "xyz":
            if r[d] == -1 and prec.isParameter(at_level):
                r[d] = prec.get(at_level, r[d])
            if r[d] == -1:
                r[d] = prec.get("Retain Nodes", r[d])
        self.rx, self.ry, self.rz = r["x"], r["y"], r["z"]

        self.link_retained_nodes = prec.get("Eliminate Retained Nodes Together", True)
        self.link_velocities = prec.get("Eliminate Velocities Together", True)

        if prob.isParameter("Equations"):
            eqn = prob.get("Equations", "Undefined Problem")
            is_complex = prob.get("Complex Arithmetic", False)
            factor = 2 if is_complex else 1
            if eqn == "Laplace":
                if not is_complex:
                    prob.get("Degrees of Freedom", 1)
                    prob.sublist("Variable 0").get("Variable Type", "Laplace")
                else:
                    prob.get("Degrees of Freedom", 2)
                    prob.sublist("Variable 0").get("Variable Type", "Laplace")
                    prob.sublist("Variable 1").get("Variable Type", "Laplace")
            elif eqn.startswith("Stokes") or eqn == "Bous-C":
                if eqn == "Bous-C":
                    prob.get("Degrees of Freedom", self.dim + 2)
                    pvar = prob.get("Pressure Variable", self.dim + 1)
                else:
                    prob.get("Degrees of Freedom", self.dim + 1)
                    pvar = prob.get("Pressure Variable", self.dim)
                dof = prob.get("Degrees of Freedom", 1)
                for i in range(self.dim * factor):
                    prob.sublist("Variable %d" % i).get("Variable Type", "Velocity")
                for i in range(pvar * factor, pvar * factor + factor):
                    prob.sublist("Variable %d" % i).get("Variable Type", "Pressure")
                for i in range(dof):
                    if not prob.isSublist("Variable %d" % i):
                        prob.sublist("Variable %d" % i).get("Variable Type", "Laplace")
                if eqn in ("Stokes-B", "Stokes-L", "Stokes-T"):
                    if is_complex:
                        raise ValueError("complex Stokes-B not implemented")
                    prob.get("Retained Pressure Nodes", 2)
                    if prec.get("Fix Pressure Level", True):
                        prec.get("Fix GID 1", factor * pvar)
                        prec.get("Fix GID 2", factor * dof + factor * pvar)
                else:
                    if prec.get("Fix Pressure Level", True):
                        prec.get("Fix GID 1", factor * pvar)
                        if is_complex:
                            prec.get("Fix GID 2", factor * pvar + 1)
                    prob.get("Retained Pressure Nodes", 1)
            else:
                raise ValueError("'Equations' parameter not recognized")
        if not prob.isParameter("Degrees of Freedom"):
            raise ValueError("the 'Problem' sublist must contain 'Degrees of Freedom'")
        self.dof = prob.get("Degrees of Freedom", 1)
        self.retain_pressures = prob.get("Retained Pressure Nodes", 1)

        self.variable_type = [None] * self.dof
        pcount = vcount = 0
        for i in range(self.dof):
            vt = prob.sublist("Variable %d" % i).get("Variable Type", "Laplace")
            if vt == "Laplace":
                self.variable_type[i] = V_V
            elif vt == "Velocity U" or (vt == "Velocity" and vcount == 0):
                self.variable_type[i] = V_U
                vcount += 1
            elif vt == "Velocity V" or (vt == "Velocity" and vcount == 1):
                self.variable_type[i] = V_V
                vcount += 1
            elif vt == "Velocity W" or (vt == "Velocity" and vcount == 2):
                self.variable_type[i] = V_W
                vcount += 1
            elif vt == "Pressure":
                pvar = i
                self.variable_type[i] = PRESSURE
                pcount += 1
            elif vt == "Interior":
                self.variable_type[i] = INTERIOR
            else:
                raise ValueError("Variable type %s does not exist" % vt)
        if pcount > 1:
            raise ValueError("Can only have one 'Pressure' variable")
        prob.get("Pressure Variable", pvar)
        self.pvar = pvar
        self.bgrid_transform = prec.get("B-Grid Transform", False)
        # NON-REFERENCE extension (default off = reference behaviour), see DESIGN.md:
        # separator ("tube") pressure groups get the velocities' link type, so their
        # non-V-sum nodes join the edge-velocity block instead of forming an all-zero block.
        self.link_tube_pressures = prec.get("Eliminate Tube Pressures With Velocities", False)

    # -- src/HYMLS_BasePartitioner.cpp:321-346 --------------------------------
    def set_next_level_parameters(self, params):
        prec = params.sublist("Preconditioner")
        nsx, nsy, nsz = self.sx * self.cx, self.sy * self.cy, self.sz * self.cz
        if prec.isParameter("Separator Length (x)"):
            prec.set("Separator Length (x)", nsx)
            prec.set("Separator Length (y)", nsy)
            prec.set("Separator Length (z)", nsz)
        else:
            prec.set("Separator Length", nsx)
        if prec.isParameter("Coarsening Factor (x)"):
            prec.set("Coarsening Factor (x)", self.cx)
            prec.set("Coarsening Factor (y)", self.cy)
            prec.set("Coarsening Factor (z)", self.cz)
        else:
            prec.set("Coarsening Factor", self.cx)

    # -- src/HYMLS_CartesianPartitioner.cpp:80-121 ----------------------------
    def subdomain_position(self, sd, sx, sy, sz):
        npx = (self.nx - 1) // sx + 1
        npy = (self.ny - 1) // sy + 1
        npz = (self.nz - 1) // sz + 1
        return (sd % npx) * sx, ((sd // npx) % npy) * sy, ((sd // npx // npy) % npz) * sz

    def subdomain_id(self, sx, sy, sz, x, y, z):
        npx = (self.nx - 1) // sx + 1
        npy = (self.ny - 1) // sy + 1
        return (z // sz * npy + y // sy) * npx + x // sx

    def num_global_parts(self, sx=None, sy=None, sz=None):
        sx = self.sx if sx is None else sx
        sy = self.sy if sy is None else sy
        sz = self.sz if sz is None else sz
        return ((self.nx - 1) // sx + 1) * ((self.ny - 1) // sy + 1) * ((self.nz - 1) // sz + 1)

    def __call__(self, i, j, k):
        return self.subdomain_id(self.sx, self.sy, self.sz, i, j, k)

    def ind2sub(self, gid):
        # src/HYMLS_Tools.cpp:662-680
        rem = gid
        var = rem % self.dof
        rem //= self.dof
        i = rem % self.nx
        rem //= self.nx
        j = rem % self.ny
        rem //= self.ny
        k = rem % self.nz
        return i, j, k, var

    def pid(self, gid):
        i, j, k, _ = self.ind2sub(gid)
        return self.pid_map[self.subdomain_id(self.sx, self.sy, self.sz, i, j, k)]

    # -- src/HYMLS_BasePartitioner.cpp:361-586 --------------------------------
    def create_pid_map(self):
        nx, ny, nz = self.nx, self.ny, self.nz
        sx, sy, sz = self.sx, self.sy, self.sz
        nparts = self.num_global_parts(sx, sy, sz)
        P = self.nprocs_comm
        if P == 1 or nparts == 1:
            self.nprocs = 1
            self.pid_map = [0] * nparts
            return
        pid_map = [-1] * nparts
        pid_groups = [[] for _ in range(nparts)]
        sd_pid_num = [0] * nparts
        cx = find_coarsening_factor(self.cx)
        cy = find_coarsening_factor(self.cy)
        cz = find_coarsening_factor(self.cz)
        while sx < nx or sy < ny or sz < nz:
            sx *= cx
            sy *= cy
            if nz > 1:
                sz *= cz
        sx2, sy2, sz2 = sx, sy, sz
        nprocs = 0

        def wrap(x, y, z):
            return (x % nx + nx) % nx, (y % ny + ny) % ny, (z % nz + nz) % nz

        for _ in range(1000):
            nparts = self.num_global_parts(sx, sy, sz)
            prev_nprocs = nprocs
            prev_groups = [list(g) for g in pid_groups]
            for i in range(nparts):
                x, y, z = wrap(*self.subdomain_position(i, sx, sy, sz))
                sd = self.subdomain_id(self.sx, self.sy, self.sz, x, y, z)
                if len(pid_groups[sd]) == 0:
                    pid_groups[sd].append(nprocs)
                    nprocs += 1
            if nprocs > P:
                nprocs = prev_nprocs
                pid_groups = prev_groups
                break
            sx2, sy2, sz2 = sx, sy, sz
            sx //= cx
            sy //= cy
            if nz > 1:
                sz //= cz
            if sx < self.sx or sy < self.sy or sz < self.sz:
                sx, sy, sz = sx2, sy2, sz2
                break

        nparts = self.num_global_parts()
        for _ in range(1000):
            if nprocs >= P:
                break
            for sd in range(nparts):
                if nprocs >= P:
                    break
                if len(pid_groups[sd]) != 0:
                    pid_groups[sd].append(nprocs)
                    nprocs += 1

        nparts = self.num_global_parts(sx, sy, sz)
        for i in range(nparts):
            x, y, z = wrap(*self.subdomain_position(i, sx, sy, sz))
            sd = self.subdomain_id(self.sx, self.sy, self.sz, x, y, z)
            if pid_map[sd] != -1:
                continue
            sd2 = self.subdomain_id(sx2, sy2, sz2, x, y, z)
            x, y, z = wrap(*self.subdomain_position(sd2, sx2, sy2, sz2))
            sd2 = self.subdomain_id(self.sx, self.sy, self.sz, x, y, z)
            if len(pid_groups[sd2]) == 0:
                raise RuntimeError("Invalid subdomain index %d" % sd)
            pid_map[sd] = pid_groups[sd2][sd_pid_num[sd2] % len(pid_groups[sd2])]
            sd_pid_num[sd2] += 1

        nparts = self.num_global_parts()
        for i in range(nparts):
            if pid_map[i] == -1:
                x, y, z = wrap(*self.subdomain_position(i, self.sx, self.sy, self.sz))
                sd = self.subdomain_id(self.sx, self.sy, self.sz, x, y, z)
                if pid_map[sd] != -1:
                    pid_map[i] = pid_map[sd]
                    continue
                sd = self.subdomain_id(sx, sy, sz, x, y, z)
                x, y, z = wrap(*self.subdomain_position(sd, sx, sy, sz))
                sd = self.subdomain_id(self.sx, self.sy, self.sz, x, y, z)
                if pid_map[sd] == -1:
                    raise RuntimeError("Invalid subdomain index %d" % sd)
                pid_map[i] = pid_map[sd]
        self.pid_map = pid_map
        self.nprocs = len(set(pid_map))

    # -- src/HYMLS_CartesianPartitioner.cpp:123-222 ---------------------------
    def partition(self):
        self.create_pid_map()
        self.sd_map = [sd for sd in range(self.num_global_parts())
                       if self.pid_map[sd] == self.mypid]
        return self

    def num_local_parts(self):
        return len(self.sd_map)

    def owned_gids(self):
        """GIDs of the repartitioned (cartesian) map on this rank, ascending
        (src/HYMLS_BasePartitioner.cpp:686-775 sorts them)."""
        n = self.nx * self.ny * self.nz
        cell = np.arange(n, dtype=np.int64)
        i = cell % self.nx
        j = (cell // self.nx) % self.ny
        k = cell // (self.nx * self.ny)
        npx = (self.nx - 1) // self.sx + 1
        npy = (self.ny - 1) // self.sy + 1
        sd = (k // self.sz * npy + j // self.sy) * npx + i // self.sx
        mine = np.asarray(self.pid_map)[sd] == self.mypid
        cells = cell[mine]
        return (cells[:, None] * self.dof + np.arange(self.dof)[None, :]).reshape(-1)

    # -- src/HYMLS_CartesianPartitioner.cpp:224-263 ---------------------------
    @staticmethod
    def _start_and_end(pos, idx, idx_max, dim, mx, perio):
        ln = max((mx + idx_max - 1) // idx_max, 1)
        if idx == idx_max:
            typ = 2
        elif idx >= 0:
            typ = 1
        else:
            typ = 0
        start = idx
        if idx == idx_max:
            start = mx
        elif idx > 0:
            start = min(ln * idx, mx)
        end = start + 1
        if typ == 1:
            end = min(ln * (idx + 1), mx)
        if not perio:
            if pos == 0 and idx == -1:
                return True, typ, start, end
            if pos + mx + 1 == dim:
                if idx == idx_max:
                    return True, typ, start, end
                if idx == idx_max - 1:
                    end += 1
        if start == end:
            return True, typ, start, end
        return False, typ, start, end

    # -- src/HYMLS_CartesianPartitioner.cpp:265-408 ---------------------------
    def get_groups(self, sd_local):
        """returns (interior gids [unsorted emission order], [(type, [gids])...])"""
        nx, ny, nz, dof = self.nx, self.ny, self.nz, self.dof
        gsd = self.sd_map[sd_local]
        xpos, ypos, zpos = self.subdomain_position(gsd, self.sx, self.sy, self.sz)
        xmax = min(nx - xpos - 1, self.sx - 1)
        ymax = min(ny - ypos - 1, self.sy - 1)
        zmax = min(nz - zpos - 1, self.sz - 1)
        if xmax == 0 or ymax == 0 or (zmax == 0 and nz > 1):
            raise ValueError("Can't have a subdomain of size 1")
        iidx_max = self.rx if self.rx > 1 else 1
        jidx_max = self.ry if self.ry > 1 else 1
        kidx_max = self.rz if self.rz > 1 else 1
        interior = []
        groups = []  # list of [type, nodes]
        retained = []
        for kidx in range(-1, kidx_max + 1):
            kint = 0 <= kidx < kidx_max
            skip, ktype, kstart, kend = self._start_and_end(
                zpos, kidx, kidx_max, nz, zmax, self.perio & Z_PERIO)
            if skip:
                continue
            for jidx in range(-1, jidx_max + 1):
                jint = 0 <= jidx < jidx_max
                skip, jtype, jstart, jend = self._start_and_end(
                    ypos, jidx, jidx_max, ny, ymax, self.perio & Y_PERIO)
                if skip:
                    continue
                for iidx in range(-1, iidx_max + 1):
                    iint = 0 <= iidx < iidx_max
                    skip, itype, istart, iend = self._start_and_end(
                        xpos, iidx, iidx_max, nx, xmax, self.perio & X_PERIO)
                    if skip:
                        continue
                    for d in range(dof):
                        vt = self.variable_type[d]
                        nodes2 = None
                        if vt in (PRESSURE, INTERIOR) and (iidx == -1 or jidx == -1 or kidx == -1):
                            continue
                        elif ((iint and jint and kint) or vt == INTERIOR or
                              (vt == PRESSURE and ((iint and jint) or (iint and kint) or
                                                   (jint and kint) or self.retain_pressures > 1))):
                            nodes = interior
                        else:
                            typ = -1000
                            if self.link_retained_nodes:
                                typ = 2 * dof * (itype + 3 * (jtype + 3 * ktype))
                            if not ((self.link_velocities and vt in (V_U, V_V, V_W)) or
                                    (self.link_tube_pressures and vt == PRESSURE)):
                                typ += 2 * d
                            grp = [typ, []]
                            groups.append(grp)
                            nodes = grp[1]
                            if self.bgrid_transform:
                                grp2 = [typ + 1, []]
                                groups.append(grp2)
                                nodes2 = grp2[1]
                        for k in range(kstart, kend):
                            for j in range(jstart, jend):
                                for i in range(istart, iend):
                                    gid = (d + ((i + xpos + nx) % nx) * dof +
                                           ((j + ypos + ny) % ny) * nx * dof +
                                           ((k + zpos + nz) % nz) * nx * ny * dof)
                                    if (vt == PRESSURE and i >= 0 and j >= 0 and k >= 0 and
                                            len(retained) < self.retain_pressures):
                                        retained.append(gid)
                                    elif nodes2 is not None and (i + xpos + j + ypos) % 2:
                                        nodes2.append(gid)
                                    else:
                                        nodes.append(gid)
        groups = [g for g in groups if len(g[1]) > 0]
        for gid in retained:
            groups.append([-1, [gid]])
        return interior, [(g[0], g[1]) for g in groups]


class HierarchicalMap:
    """Single-rank-view restatement of HierarchicalMap after FillComplete.

    interior[sd]            sorted GIDs
    groups[sd]              list of (type, sorted GIDs), empty ones removed
    unique_groups[sd]       groups whose first GID was not seen in an earlier local sd
    linked[sd]              list of lists of group indices (into groups[sd]) of equal type>=0
    """

    def __init__(self, interior, groups, base_gids, overlapping_gids=None, owned=None):
        # FillComplete, src/HYMLS_HierarchicalMap.cpp:144-285
        present = set(int(g) for g in (overlapping_gids if overlapping_gids is not None else base_gids))
        # On one rank every GID handed out by the partitioner that exists in the
        # base map is "present on its owner"; the Import at :216-218 reduces to that.
        self.base_gids = np.asarray(base_gids, dtype=np.int64)
        self.owned = set(int(g) for g in (owned if owned is not None else base_gids))
        self.interior = []
        self.groups = []
        for sd in range(len(interior)):
            self.interior.append([g for g in interior[sd] if g in present])
            grp = []
            for typ, nodes in groups[sd]:
                kept = [g for g in nodes if g in present]
                if kept:
                    grp.append((typ, kept))
            self.groups.append(grp)
        self.unique_groups = []
        seen = set()
        all_gids = []
        for sd in range(len(interior)):
            all_gids.extend(self.interior[sd])
            uniq = []
            for gi, (typ, nodes) in enumerate(self.groups[sd]):
                if nodes[0] not in seen:
                    seen.add(nodes[0])
                    uniq.append(gi)
                    all_gids.extend(nodes)
            self.unique_groups.append(uniq)
        self.overlapping_map = np.asarray(all_gids, dtype=np.int64)
        self.linked = [self._link([g for g in range(len(self.groups[sd]))], sd)
                       for sd in range(len(interior))]

    def _link(self, group_ids, sd):
        # LinkSeparators, src/HYMLS_HierarchicalMap.cpp:120-142
        out = []
        for gi in group_ids:
            typ = self.groups[sd][gi][0]
            found = False
            if typ >= 0:
                for lg in out:
                    if self.groups[sd][lg[0]][0] == typ:
                        lg.append(gi)
                        found = True
                        break
            if not found:
                out.append([gi])
        return out

    def num_subdomains(self):
        return len(self.interior)

    # SpawnInterior :436-466
    def interior_map(self):
        return np.asarray([g for sd in self.interior for g in sd], dtype=np.int64)

    # SpawnSeparators :470-508 (overlapping = all unique groups, map = owned ones)
    def separator_overlapping_map(self):
        return np.asarray([g for sd in range(self.num_subdomains())
                           for gi in self.unique_groups[sd]
                           for g in self.groups[sd][gi][1]], dtype=np.int64)

    def separator_map(self):
        return np.asarray([g for g in self.separator_overlapping_map() if int(g) in self.owned],
                          dtype=np.int64)

    # SpawnLocalSeparators :512-542
    def local_groups(self, sd):
        """group indices (into groups[sd]) of unique groups whose first GID is owned"""
        return [gi for gi in self.unique_groups[sd] if self.groups[sd][gi][1][0] in self.owned]

    def local_linked(self, sd):
        return self._link(self.local_groups(sd), sd)


class OverlappingPartitioner(HierarchicalMap):
    """src/HYMLS_OverlappingPartitioner.cpp:31-159 (Cartesian only)."""

    def __init__(self, params, level=0, base_gids=None, overlapping_gids=None,
                 nprocs=1, mypid=0):
        self.params = params
        self.level = level
        method = params.sublist("Preconditioner").get("Partitioner", "Cartesian")
        if method == "Cartesian":
            part = CartesianPartitioner(params, level, nprocs, mypid).partition()
        elif method == "Skew Cartesian":
            from .skew import SkewCartesianPartitioner
            part = SkewCartesianPartitioner(params, level, nprocs, mypid).partition()
        else:
            raise NotImplementedError("Up to now we only support Cartesian partitioning")
        self.partitioner = part
        self.next_level_params = params.copy()
        part.set_next_level_parameters(self.next_level_params)
        owned = part.owned_gids()
        if base_gids is None:
            base = owned
        else:
            # level >= 1: the map handed in is restricted to GIDs this rank owns
            own = set(int(g) for g in owned)
            base = np.asarray([g for g in base_gids if int(g) in own], dtype=np.int64) \
                if nprocs > 1 else np.asarray(base_gids, dtype=np.int64)
        interior, groups = [], []
        for sd in range(part.num_local_parts()):
            it, gr = part.get_groups(sd)
            interior.append(sorted(it))
            groups.append([(t, sorted(n)) for t, n in gr])
        if nprocs == 1:
            present = base if overlapping_gids is None else overlapping_gids
            HierarchicalMap.__init__(self, interior, groups, base, present, owned=base)
        else:
            # level 0 only for fake multi-rank views: every GID of the grid exists somewhere
            n = part.nx * part.ny * part.nz * part.dof
            allg = np.arange(n, dtype=np.int64) if base_gids is None else overlapping_gids
            HierarchicalMap.__init__(self, interior, groups, base, allg, owned=base)

    def spawn_next_level(self, vsum_gids, overlapping_vsum_gids):
        return OverlappingPartitioner(self.next_level_params.copy(), self.level + 1,
                                      base_gids=vsum_gids,
                                      overlapping_gids=overlapping_vsum_gids)
