"""Skew Cartesian partitioner oracle (ORACLE -- test infrastructure only).

Restates src/HYMLS_SkewCartesianPartitioner.cpp (45-degree rotated, octahedral subdomains):
  buildPlane45 :27-78, GetSubdomainPosition :128-160, GetSubdomainID :162-207, NumGlobalParts :219-238,
  CreateSubdomainMap :240-272, Partition :276-368, getTemplate :372-565, solveGroups :567-654,
  GetGroups :656-812.
Integer arithmetic follows the C++ semantics (all divisions below act on non-negative operands or are
exact, so Python's floor division agrees with C++ truncation).
"""
import numpy as np

from .partitioner import (CartesianPartitioner, PRESSURE, V_U, V_V, V_W, X_PERIO, Y_PERIO, Z_PERIO)


def build_plane45(first, length, dir_x, dir_y, typ):
    left = right = first
    height = 2 * length
    extra = False
    dir1 = dir_y + dir_x
    dir2 = dir_y - dir_x
    if typ == 0:
        left -= dir_x
        height += 1
        extra = True
    elif typ == 3:
        height += 1
        extra = True
    ptr = [0]
    plane = []
    for i in range(height - 1):
        j = left
        while j <= right:
            plane.append(j)
            j += dir_x
        ptr.append(len(plane))
        if i < length - 1:
            left += dir2
            right += dir1
        elif extra and i == length - 1:
            left += dir_y
            right += dir_y
        else:
            left += dir1
            right += dir2
    return ptr, plane


class SkewCartesianPartitioner(CartesianPartitioner):
    # -- :128-160 ---------------------------------------------------------------------------------
    def subdomain_position_status(self, sd, sx, sy, sz):
        npx = self.nx // sx
        npy = self.ny // sy
        tot2d = npx * npy
        per_layer = 2 * tot2d + npx + npy
        per_row = 2 * npx + 1
        Z = sd // per_layer if per_layer > 0 else 0
        Y = ((sd - Z * per_layer) // per_row) * 2 - 1
        X = ((sd - Z * per_layer) % per_row) * 2
        if X >= npx * 2:
            X -= npx * 2 + 1
            Y += 1
        x = (X * sx) // 2
        y = (Y * sx) // 2 + sx // 2
        z = Z * sx
        status = 0
        if x == self.nx - sx // 2 and self.perio & X_PERIO:
            status = 1
        if y == self.ny and self.perio & Y_PERIO:
            status = 1
        if z == self.nz and self.perio & Z_PERIO:
            status = 1
        return x, y, z, status

    def subdomain_position(self, sd, sx, sy, sz):
        return self.subdomain_position_status(sd, sx, sy, sz)[:3]

    # -- :162-207 ---------------------------------------------------------------------------------
    def subdomain_id(self, sx, sy, sz, x, y, z):
        npx = self.nx // sx
        npy = self.ny // sy
        npz = self.nz // sz
        dir1 = npx + 1
        dir2 = npx
        dir3 = 2 * npx * npy + npx + npy
        xc, yc, zc = x // sx, y // sx, z // sx
        sd = zc * dir3 + yc * (dir2 + dir1) + xc
        x -= xc * sx - 1
        y -= yc * sx
        z -= zc * sx
        front = y < sx - x
        right = y < x
        below = z <= y - x
        if right:
            below = z <= sx + y - x
        if not front:
            sd += dir1
        if not right:
            sd += dir2
        if not below:
            sd += dir3
        if (not front) and right and (self.perio & X_PERIO) and xc == npx - 1:
            sd -= dir2
        if (not front) and (not right) and (self.perio & Y_PERIO) and yc == npy - 1:
            sd -= dir3 - dir2
        if (not below) and (self.perio & Z_PERIO) and zc == npz - 1:
            sd -= npz * dir3
        return sd

    # -- :219-238 ---------------------------------------------------------------------------------
    def num_global_parts(self, sx=None, sy=None, sz=None):
        sx = self.sx if sx is None else sx
        sy = self.sy if sy is None else sy
        sz = self.sz if sz is None else sz
        npx, npy, npz = self.nx // sx, self.ny // sy, self.nz // sz
        per_layer = 2 * npx * npy + npx + npy
        n = per_layer
        if self.nz > 1:
            n += per_layer * npz
        return max(n, 1)

    # -- :240-368 ---------------------------------------------------------------------------------
    def partition(self):
        if self.sx != self.sy or (self.nz > 1 and self.sx != self.sz):
            raise ValueError("sx, sy and sz should be the same")
        if self.sx % 2:
            raise ValueError("sx should be even")
        if self.nx % self.sx or self.ny % self.sy or self.nz % self.sz:
            raise ValueError("domain size must be a multiple of the subdomain size")
        self.create_pid_map()
        self.sd_map = []
        for sd in range(self.num_global_parts()):
            i, j, k, status = self.subdomain_position_status(sd, self.sx, self.sy, self.sz)
            if status == 1:
                continue
            i = (i % self.nx + self.nx) % self.nx
            j = (j % self.ny + self.ny) % self.ny
            k = (k % self.nz + self.nz) % self.nz
            if self.pid_map[self.subdomain_id(self.sx, self.sy, self.sz, i, j, k)] == self.mypid:
                self.sd_map.append(sd)
        self._template()
        self._solve_groups()
        return self

    def owned_gids(self):
        n = self.nx * self.ny * self.nz
        out = []
        for cell in range(n):
            i = cell % self.nx
            j = (cell // self.nx) % self.ny
            k = cell // (self.nx * self.ny)
            if self.pid_map[self.subdomain_id(self.sx, self.sy, self.sz, i, j, k)] == self.mypid:
                out.extend(range(cell * self.dof, (cell + 1) * self.dof))
        return np.asarray(out, dtype=np.int64)

    # -- getTemplate :372-565 ---------------------------------------------------------------------
    def _template(self):
        sx, dof = self.sx, self.dof
        nx = sx * 4
        dir_x, dir_y, dir_z = dof, dof * nx, dof * nx * nx
        first = [dof * sx // 2 + dir_y + dir_z * sx,
                 dof * sx // 2 + dir_z * sx,
                 dof * sx // 2 + dir_y + dir_z * sx,
                 dof * sx // 2 + dir_y + dir_z * sx]
        base_len = [sx // 2, sx // 2 + 1, sx // 2 + 1, sx // 2]
        type_arr = [V_U, V_V, V_W, PRESSURE]
        nodes = []
        for typ in range(4):
            layers = [[] for _ in range(2 * sx + 1)]
            nodes.append(layers)
            ptr, plane = build_plane45(first[typ], base_len[typ], dir_x, dir_y, typ)
            layers[sx] = list(plane)
            if self.nz <= 1:
                continue
            bottom = []
            top = list(plane)
            row_len = [ptr[i + 1] - ptr[i] - 1 for i in range(len(ptr) - 1)]
            active = list(range(base_len[typ]))
            offset = [row_len[i] for i in active]
            for i in range(sx):
                for j in range(len(active)):
                    val = plane[ptr[active[j]] + offset[j]]
                    bottom.append(val)
                    top = [t for t in top if t != val]
                if type_arr[typ] == V_W:
                    if i % 2 == 1:
                        layers[sx + i].extend(j + i * dir_z - dir_y for j in top)
                        layers[sx + 1 + i].extend(j + (i + 1) * dir_z for j in top)
                    else:
                        layers[i].extend(j - (sx - i) * dir_z for j in bottom)
                        if i > 0:
                            layers[i - 1].extend(j - (sx - i + 1) * dir_z - dir_y for j in bottom)
                        else:
                            layers[sx - 1].extend(j - dir_z - dir_y for j in plane)
                else:
                    is_p = 1 if type_arr[typ] == PRESSURE else 0
                    if i < sx - is_p:
                        layers[i + is_p].extend(j - (sx - i - is_p) * dir_z for j in bottom)
                    layers[sx + 1 + i].extend(j + (i + 1) * dir_z for j in top)
                if i < sx - 1:
                    offset = [d - 1 for d in offset]
                    if type_arr[typ] == PRESSURE:
                        if offset[0] < 0:
                            active.append(active[-1] + 1)
                            active.pop(0)
                            offset.append(row_len[active[-1]])
                            offset.pop(0)
                    else:
                        if offset[0] < 0:
                            active.pop(0)
                            offset.pop(0)
                        elif offset[0] == 0:
                            active.append(active[-1] + 1)
                            offset.append(row_len[active[-1]])
        nodes[0].pop()
        nodes[0].pop(0)
        nodes[1].pop()
        nodes[1].pop(0)
        nodes[2].pop()
        nodes[3].pop()
        nodes[3].pop(0)
        template = [[]]
        for i in range(dof):
            if self.variable_type[i] == V_W:
                template[-1].extend(v + i for v in nodes[2][0])
                nodes[2].pop(0)
                break
        for j in range(2 * sx - 1):
            layer = []
            for i in range(dof):
                for typ in range(4):
                    if self.variable_type[i] == type_arr[typ]:
                        layer.extend(v + i for v in nodes[typ][j])
            layer.sort()
            template.append(layer)
        self.template = template

    # -- solveGroups :567-654 ----------------------------------------------------------------------
    def _solve_groups(self):
        sx, dof = self.sx, self.dof
        nx = sx * 4
        dir_x, dir_y, dir_z = dof * sx, dof * nx * sx, dof * nx * nx * sx
        first = dir_x + dir_y + dir_z
        d1 = (dir_y + dir_x) // 2
        d2 = (dir_y - dir_x) // 2 + dir_z
        d3 = dir_z
        positions = [0, -d3, d3, -d2, -d2 - d3, -d2 + d3, d2, d2 - d3, d2 + d3,
                     -d1, -d1 - d3, -d1 + d3, -d1 - d2, -d1 - d2 - d3, -d1 - d2 + d3, -d1 + d2,
                     -d1 + d2 - d3, -d1 + d2 + d3, d1, d1 - d3, d1 + d3, d1 - d2,
                     d1 - d2 - d3, d1 - d2 + d3, d1 + d2, d1 + d2 - d3, d1 + d2 + d3]
        temp = [v + first for layer in self.template for v in layer]
        temp_set = set(temp)
        groups = [[]]
        domains = [1]
        for node in temp:
            bits = 0
            for i, p in enumerate(positions):
                if node - p in temp_set:
                    bits += 1 << i
            for gi, dmn in enumerate(domains):
                if dmn == bits:
                    groups[gi].append(node)
                    break
            else:
                groups.append([node])
                domains.append(bits)
        out = [[groups[0]]]
        for grp in groups[1:]:
            cat = [[] for _ in range(dof)]
            for node in grp:
                cat[node % dof].append(node)
            out.append(cat)
        self.groups_template = out

    # -- GetGroups :656-812 -------------------------------------------------------------------------
    def get_groups(self, sd_local):
        dof, sx = self.dof, self.sx
        gsd = self.sd_map[sd_local]
        sdx, sdy, sdz = self.subdomain_position(gsd, self.sx, self.sy, self.sz)
        nx = 4 * sx
        groups = []
        for cat in self.groups_template:
            moved = []
            for group in cat:
                g = []
                for node in group:
                    var = node % dof
                    x = (node // dof) % nx + sdx - 1 - sx
                    y = (node // dof // nx) % nx + sdy - 1 - 3 * sx // 2
                    z = node // dof // nx // nx + sdz - 2 * sx
                    if self.perio & X_PERIO:
                        x = (x + self.nx) % self.nx
                    if self.perio & Y_PERIO:
                        y = (y + self.ny) % self.ny
                    if self.perio & Z_PERIO:
                        z = (z + self.nz) % self.nz
                    if 0 <= x < self.nx and 0 <= y < self.ny and 0 <= z < self.nz:
                        g.append(x * dof + self.nx * y * dof + self.nx * self.ny * z * dof + var)
                moved.append(g)
            groups.append(moved)
        # retained pressure nodes: first pressure nodes of the interior become groups of their own
        # (the reference erases from the vector it is iterating over: the element that slides into the
        #  erased slot is skipped -- reproduced here)
        retained = 0
        it = 0
        interior0 = groups[0][0]
        while it < len(interior0):
            node = interior0[it]
            if self.variable_type[node % dof] == PRESSURE:
                groups.append([[node]])
                interior0.pop(it)
                retained += 1
                if retained >= self.retain_pressures:
                    break
            it += 1
        interior = list(groups[0][0])
        seps = []
        typ = 1
        for i in range(1, len(groups)):
            typ += 1
            for group in groups[i]:
                new = {}
                for node in group:
                    cell = node // dof
                    owner = self.subdomain_id(self.sx, self.sy, self.sz, cell % self.nx,
                                              (cell // self.nx) % self.ny, cell // (self.nx * self.ny))
                    if owner in new:
                        new[owner][1].append(node)
                    else:
                        new[owner] = [typ if self.link_velocities else -1, [node]]
                for owner in sorted(new):
                    t, nodes = new[owner]
                    if self.rx > 1:
                        if not self.link_velocities:
                            typ += 1
                        ln = len(nodes)
                        new_len = max((ln + self.rx - 1) // self.rx, 1)
                        parts = (ln - 1) // new_len + 1
                        for j in range(parts):
                            t2 = typ if (self.link_velocities or self.link_retained_nodes) else -1
                            seps.append([t2, nodes[j * new_len:min((j + 1) * new_len, ln)]])
                    else:
                        seps.append([t, nodes])
        # velocity nodes on the (non-periodic) far boundaries are Dirichlet rows: not separators
        for grp in seps:
            for node in list(grp[1]):
                x = (node // dof) % self.nx
                y = (node // dof // self.nx) % self.ny
                z = node // dof // self.nx // self.ny
                vt = self.variable_type[node % dof]
                hit = ((dof > 1 and x == self.nx - 1 and vt == V_U and not (self.perio & X_PERIO)) or
                       (dof > 1 and y == self.ny - 1 and vt == V_V and not (self.perio & Y_PERIO)) or
                       (self.nz > 1 and dof > 1 and z == self.nz - 1 and vt == V_W and not (self.perio & Z_PERIO)))
                if hit:
                    if self.subdomain_id(self.sx, self.sy, self.sz, x, y, z) == gsd:
                        interior.append(node)
                    grp[1].remove(node)
        interior.sort()
        return interior, [(g[0], g[1]) for g in seps]
