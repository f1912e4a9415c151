"""Pins the index-set oracle to the closed-form expectations of the reference's own unit tests
(testSuite/unit_tests/HYMLS_CartesianPartitioner.cpp, HYMLS_OverlappingPartitioner.cpp)."""
import numpy as np
import pytest

from oracle.params import ParameterList
from oracle.partitioner import CartesianPartitioner, OverlappingPartitioner
from tests.common import stokes_var_params


def _part_params(nx, ny, nz, dof, sx, sy, sz):
    p = ParameterList()
    pr = p.sublist("Problem")
    pr.set("nx", nx); pr.set("ny", ny); pr.set("nz", nz); pr.set("Degrees of Freedom", dof)
    pc = p.sublist("Preconditioner")
    pc.set("Separator Length (x)", sx); pc.set("Separator Length (y)", sy); pc.set("Separator Length (z)", sz)
    return p


def test_partition2d_ids():  # unit_tests/HYMLS_CartesianPartitioner.cpp:16-40
    part = CartesianPartitioner(_part_params(8, 8, 1, 4, 4, 2, 1)).partition()
    assert part(0, 0, 0) == 0 and part(0, 3, 0) == 2 and part(6, 3, 0) == 3


def test_partition3d_ids():  # :42-67
    part = CartesianPartitioner(_part_params(8, 8, 2, 4, 4, 2, 2)).partition()
    assert (part(0, 0, 0), part(0, 3, 0), part(6, 3, 0), part(0, 3, 1)) == (0, 2, 3, 2)


def test_5dof_nodes_all_covered():  # :69-106 (Bous-C, dof 5)
    p = ParameterList()
    pr = p.sublist("Problem")
    pr.set("nx", 8); pr.set("ny", 8); pr.set("nz", 8); pr.set("Equations", "Bous-C")
    p.sublist("Preconditioner").set("Separator Length", 4)
    part = CartesianPartitioner(p).partition()
    n = 8 * 8 * 8 * 5
    seen = np.zeros(n, dtype=bool)
    for sd in range(part.num_local_parts()):
        it, gr = part.get_groups(sd)
        seen[it] = True
        for _, nodes in gr:
            seen[nodes] = True
    assert seen.all()


def test_same_part_every_proc():  # :108-133: 64 fake ranks on 32^3, 8 subdomains each
    for pid in range(0, 64, 7):
        p = ParameterList()
        pr = p.sublist("Problem")
        pr.set("nx", 32); pr.set("ny", 32); pr.set("nz", 32); pr.set("Equations", "Stokes-C")
        p.sublist("Preconditioner").set("Separator Length", 4)
        part = CartesianPartitioner(p, nprocs=64, mypid=pid).partition()
        assert part.num_local_parts() == 8 * 8 * 8 // 64


def test_pid_map_8_ranks_is_2x2x2_bricks():
    """16^3 subdomains on 8 ranks -> 2x2x2 bricks of 8^3 subdomains."""
    p = ParameterList()
    pr = p.sublist("Problem")
    pr.set("nx", 64); pr.set("ny", 64); pr.set("nz", 64); pr.set("Equations", "Stokes-C")
    p.sublist("Preconditioner").set("Separator Length", 4)
    part = CartesianPartitioner(p, nprocs=8, mypid=0).partition()
    pm = np.asarray(part.pid_map).reshape(16, 16, 16)
    assert part.nprocs == 8
    for r in range(8):
        zz, yy, xx = np.where(pm == r)
        assert len(zz) == 512
        assert xx.max() - xx.min() == 7 and yy.max() - yy.min() == 7 and zz.max() - zz.min() == 7


@pytest.mark.parametrize("nx,ny,sx,sy", [(8, 8, 4, 4), (16, 16, 4, 4), (16, 8, 4, 4), (64, 64, 16, 16)])
def test_stokes2d_groups(nx, ny, sx, sy):  # unit_tests/HYMLS_OverlappingPartitioner.cpp:341-535
    dof = 3
    op = OverlappingPartitioner(stokes_var_params(2, nx, ny, 1, sx), 0)
    nsx, nsy = nx // sx, ny // sy
    for sd in range(op.num_subdomains()):
        gsd = op.partitioner.sd_map[sd]
        substart = gsd % nsx * nx // nsx * dof + gsd // nsx * ny // nsy * dof * nx
        is_group = [1] * 9
        if (gsd + 1) % nsx == 0:
            is_group[2] = is_group[5] = is_group[8] = 0
        if gsd // nsx == nsy - 1:
            is_group[6] = is_group[7] = is_group[8] = 0
        if gsd % nsx == 0:
            is_group[0] = is_group[3] = is_group[6] = 0
        if gsd // nsx == 0:
            is_group[0] = is_group[1] = is_group[2] = 0
        num_groups = sum(is_group) * 2 - 1 + 1 + is_group[8]
        assert len(op.groups[sd]) == num_groups - 1
        interior = op.interior[sd]
        right, bottom = (gsd + 1) % nsx == 0, gsd // nsx == nsy - 1
        exp = []
        pos = 0
        for y in range(sy):
            for x in range(sx):
                for d in range(dof):
                    if right and bottom:
                        cond = True
                    elif right:
                        cond = (x < sx and y < sy - 1) or d == 2
                    elif bottom:
                        cond = (x < sx - 1 and y < sy) or d == 2
                    else:
                        cond = (x < sx - 1 and y < sy - 1) or d == 2
                    if cond and not (d == 2 and pos == 2) and not (d == 2 and x == sx - 1 and y == sy - 1):
                        exp.append(substart + x * dof + y * nx * dof + d)
                        pos += 1
        # the reference walks `exp` and compares position by position (a prefix check) ...
        assert interior[:len(exp)] == exp
        # ... and pins the length separately
        if right and bottom:
            assert len(interior) == sx * sy * dof - 1
        elif right:
            assert len(interior) == sx * (sy - 1) * 2 + sx * sy - 1
        elif bottom:
            assert len(interior) == sy * (sx - 1) * 2 + sx * sy - 1
        else:
            assert len(interior) == (sx - 1) * (sy - 1) * 2 + sx * sy - 2
        for _, nodes in op.groups[sd]:
            first = nodes[0] // dof
            if first == substart // dof - nx or first == substart // dof + nx * (sx - 1):
                assert len(nodes) == (sx if right else sx - 1)
                assert nodes == [nodes[0] + i * dof for i in range(len(nodes))]
            elif first == substart // dof + sy - 1 or first == substart // dof - 1:
                assert len(nodes) == (sy if bottom else sy - 1)
                assert nodes == [nodes[0] + i * nx * dof for i in range(len(nodes))]
            else:
                assert len(nodes) == 1


@pytest.mark.parametrize("nx,ny,nz,s", [(8, 8, 8, 4), (16, 16, 16, 4), (16, 8, 8, 4), (4, 4, 4, 2),
                                        (8, 4, 4, 4), (16, 16, 16, 8)])
def test_stokes3d_groups(nx, ny, nz, s):  # unit_tests/HYMLS_OverlappingPartitioner.cpp:537-672
    dof = 4
    op = OverlappingPartitioner(stokes_var_params(3, nx, ny, nz, s), 0)
    nsx, nsy, nsz = nx // s, ny // s, nz // s
    for sd in range(op.num_subdomains()):
        gsd = op.partitioner.sd_map[sd]
        substart = (gsd % nsx * nx // nsx * dof + (gsd % (nsx * nsy)) // nsx * ny // nsy * dof * nx +
                    gsd // (nsx * nsy) * nz // nsz * dof * nx * ny)
        g = [1] * 27
        if (gsd + 1) % nsx == 0:
            for i in range(2, 27, 3):
                g[i] = 0
        if (gsd % (nsx * nsy)) // nsx == nsy - 1:
            for i in range(3):
                for j in range(3):
                    g[6 + i + j * 9] = 0
        if gsd // (nsx * nsy) == nsz - 1:
            for i in range(18, 27):
                g[i] = 0
        if gsd % nsx == 0:
            for i in range(0, 27, 3):
                g[i] = 0
        if (gsd % (nsx * nsy)) // nsx == 0:
            for i in range(3):
                for j in range(3):
                    g[i + j * 9] = 0
        if gsd // (nsx * nsy) == 0:
            for i in range(9):
                g[i] = 0
        num_groups = sum(g) * 3 - 2 + 1 + g[17] + g[23] + g[25] + g[26]
        assert len(op.groups[sd]) == num_groups - 1
        interior = op.interior[sd]
        if g[14] == 0 and g[16] == 0 and g[22] == 0:
            assert len(interior) == s * s * s * dof - 1
            exp = []
            for i in range(len(interior) // dof + 1):
                for d in range(dof):
                    if d == 3 and len(exp) == 3:
                        continue
                    if len(exp) < len(interior):
                        exp.append(substart + (i % s) * dof + ((i // s) % s) * nx * dof +
                                   i // (s * s) * nx * ny * dof + d)
            assert interior == exp
        elif num_groups == 27 * 3 - 2 + 1 + 4:
            assert len(interior) == (s - 1) ** 3 * dof - 1 + 3 * (s - 1) ** 2
            total = len(interior) + sum(len(n) for _, n in op.groups[sd])
            assert total == s ** 3 * dof + ((s + 1) * (s + 1) + (s + 1) * s + s * s) * (dof - 1)
    # every GID appears exactly once in the overlapping (row) map
    assert sorted(op.overlapping_map.tolist()) == list(range(nx * ny * nz * dof))


def test_survey_size_classes():
    """interior sizes 134 (sx=4) and 1518 (sx=8), m_S = 305 / 1181."""
    for s, n_i, m_s in ((4, 134, 305), (8, 1518, 1181)):
        op = OverlappingPartitioner(stokes_var_params(3, 3 * s, 3 * s, 3 * s, s), 0)
        sd = 13  # centre subdomain of 3x3x3
        assert len(op.interior[sd]) == n_i
        assert sum(len(n) for _, n in op.groups[sd]) == m_s
        assert len(op.local_groups(sd)) == 26
