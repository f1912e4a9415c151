"""Bit-exact index maps: the C++ host partitioner (through the C ABI, no GPU needed) against the oracle,
which tests/test_oracle_partitioner.py pins to the reference's own unit-test expectations."""
import ctypes
import os
import re

import numpy as np
import pytest

import hymls_b200 as hb
from hymls_b200 import api
from oracle import hymls as ohymls
from oracle.partitioner import CartesianPartitioner
from tests.common import make_params


def _dictify(p):
    return {k: (_dictify(v) if isinstance(v, dict) else v) for k, v in p.items()}


def test_library_exports_every_declared_symbol():
    lib = hb.load_library()
    header = open(os.path.join(os.path.dirname(api._HERE), "include", "hymls_b200.h")).read()
    names = set(re.findall(r"\b(hymls_b200_[a-z_0-9]+)\s*\(", header))
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_compute_without_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = make_params("Laplace", 2, 8, 4, 1)
    A = hb.galeri.create_matrix("Laplace", 2, 8)
    with pytest.raises(hb.HymlsError) as e:
        hb.Preconditioner(A, _dictify(p))  # values need the device
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    P = hb.Preconditioner(A, _dictify(p), pattern_only=True)
    P.Initialize()
    with pytest.raises(hb.HymlsError) as e:
        P.Compute()
    assert e.value.code == -3


def test_unknown_preconditioner_parameter_is_rejected():
    # validateParameters, src/HYMLS_Preconditioner.cpp:126-130 (this is what makes testSuite/cavity.xml stale)
    p = _dictify(make_params("Stokes-C", 2, 8, 4, 1))
    p["Preconditioner"]["Classifier"] = "Stokes"
    with pytest.raises(hb.HymlsError) as e:
        hb.Preconditioner(None, p)
    assert "Classifier" in str(e.value)


CASES = [
    ("Laplace", 2, 32, 4, 2, None),
    ("Laplace", 3, 16, 4, 2, None),
    ("Stokes-C", 2, 32, 4, 3, 2),
    ("Stokes-C", 2, 24, 4, 1, None),      # ragged: 24 = 6 subdomains of 4
    ("Stokes-C", 3, 16, 4, 2, 2),
    ("Stokes-C", 3, 16, 8, 1, None),
    ("Stokes-C", 3, 12, 4, 1, None),
]


@pytest.mark.parametrize("eqn,dim,nx,sx,levels,cx", CASES)
def test_maps_match_oracle_on_every_level(eqn, dim, nx, sx, levels, cx):
    extra = {"Eliminate_Tube_Pressures_With_Velocities": True} if (eqn == "Stokes-C" and dim == 3) else {}
    p = make_params(eqn, dim, nx, sx, levels, cx, **extra)
    A = hb.galeri.create_matrix(eqn, dim, nx)
    if eqn == "Stokes-C":
        A = -A
    tv = hb.galeri.create_testvector(A)
    P = hb.Preconditioner(A, _dictify(p), tv, pattern_only=True)
    P.Initialize()
    O = ohymls.Preconditioner(A, p.copy(), tv)
    O.initialize()
    assert P.NumLevels() == max(levels, 1)
    lvl = O
    for l in range(P.NumLevels()):
        hid = lvl.hid
        assert P.NumMySubdomains(l) == hid.num_subdomains()
        for sd in range(hid.num_subdomains()):
            assert np.array_equal(P.GetInteriorGroup(sd, l), np.asarray(hid.interior[sd], dtype=np.int64))
            got = P.GetSeparatorGroups(sd, l)
            assert len(got) == len(hid.groups[sd])
            for (t, g), (to, go) in zip(got, hid.groups[sd]):
                assert t == to and np.array_equal(g, np.asarray(go, dtype=np.int64))
        assert np.array_equal(P.GetMap(api.MAP_OVERLAPPING, l), hid.overlapping_map)
        assert np.array_equal(P.GetMap(api.MAP_INTERIOR, l), hid.interior_map())
        assert np.array_equal(P.GetMap(api.MAP_SEPARATOR, l), hid.separator_map())
        assert np.array_equal(P.GetMap(api.MAP_VSUM, l), lvl.schur_prec.vsum_gids)
        if l + 1 < P.NumLevels():
            # the oracle builds deeper levels during compute(); do the symbolic part by hand
            sp_ = lvl.schur_prec
            import scipy.sparse as sps
            nv = len(sp_.vsum_gids)
            lvl = ohymls.Preconditioner(sps.identity(nv, format="csr"), p.copy(), np.ones(nv), l + 1,
                                        sp_.next_hid, gids=sp_.vsum_gids)
            lvl.initialize()


@pytest.mark.parametrize("nx,sx,nprocs", [(64, 4, 8), (64, 4, 2), (64, 4, 4), (32, 4, 64), (64, 8, 8), (32, 4, 3)])
def test_pid_map_matches_oracle(nx, sx, nprocs):
    p = make_params("Stokes-C", 3, nx, sx, 1)
    got = hb.pid_map(_dictify(p), nprocs)
    ref = CartesianPartitioner(p.copy(), 0, nprocs, 0).partition().pid_map
    assert np.array_equal(got, np.asarray(ref, dtype=np.int32))


SKEW_CASES = [
    ("Stokes-C", 2, 16, 4, 1, None),
    ("Stokes-C", 2, 32, 4, 3, 2),
    ("Stokes-C", 3, 8, 4, 1, None),
    ("Stokes-C", 3, 16, 4, 2, 2),
    ("Stokes-C", 3, 16, 8, 1, None),
    ("Laplace", 2, 16, 4, 2, 2),
]


@pytest.mark.parametrize("eqn,dim,nx,sx,levels,cx", SKEW_CASES)
def test_skew_maps_match_oracle_on_every_level(eqn, dim, nx, sx, levels, cx):
    """Skew Cartesian partitioner: C++ host code vs the oracle that
    tests/test_oracle_skew.py pins to the reference's goldens."""
    import scipy.sparse as sps
    p = make_params(eqn, dim, nx, sx, levels, cx, Partitioner="Skew Cartesian")
    A = hb.galeri.create_matrix(eqn, dim, nx)
    tv = hb.galeri.create_testvector(A)
    P = hb.Preconditioner(A, _dictify(p), tv, pattern_only=True)
    P.Initialize()
    lvl = ohymls.Preconditioner(A, p.copy(), tv)
    lvl.initialize()
    for l in range(P.NumLevels()):
        hid = lvl.hid
        assert P.NumMySubdomains(l) == hid.num_subdomains()
        for sd in range(hid.num_subdomains()):
            assert np.array_equal(P.GetInteriorGroup(sd, l), np.asarray(hid.interior[sd], dtype=np.int64))
            got = P.GetSeparatorGroups(sd, l)
            assert len(got) == len(hid.groups[sd])
            for (t, g), (to, go) in zip(got, hid.groups[sd]):
                assert t == to and np.array_equal(g, np.asarray(go, dtype=np.int64))
        assert np.array_equal(P.GetMap(api.MAP_SEPARATOR, l), hid.separator_map())
        assert np.array_equal(P.GetMap(api.MAP_VSUM, l), lvl.schur_prec.vsum_gids)
        if l + 1 < P.NumLevels():
            sp_ = lvl.schur_prec
            nv = len(sp_.vsum_gids)
            lvl = ohymls.Preconditioner(sps.identity(nv, format="csr"), p.copy(), np.ones(nv), l + 1,
                                        sp_.next_hid, gids=sp_.vsum_gids)
            lvl.initialize()


@pytest.mark.parametrize("nx,sx,nprocs", [(8, 4, 4), (16, 4, 8), (32, 8, 8), (16, 4, 3)])
def test_skew_pid_map_matches_oracle(nx, sx, nprocs):
    from oracle.skew import SkewCartesianPartitioner
    p = make_params("Stokes-C", 3, nx, sx, 1, Partitioner="Skew Cartesian")
    got = hb.pid_map(_dictify(p), nprocs)
    ref = SkewCartesianPartitioner(p.copy(), 0, nprocs, 0).partition().pid_map
    assert np.array_equal(got, np.asarray(ref, dtype=np.int32))


def test_domain_decomposition_decouples_interiors():
    """Tester::isDDcorrect (src/HYMLS_Tester.cpp:253-455): no matrix entry may couple the interiors of two
    different subdomains -- on every level, also for the wider Re > 0 stencils of the shipped fixtures and for
    the skew partitioner; and the leading-rows bookkeeping of the first subdomain solve is consistent."""
    import scipy.sparse as sp
    from tests.conftest import load_fixture
    cases = [("cavity2d_32_Re1000", make_params("Stokes-C", 2, 32, 4, 3, 2)),
             ("cavity2d_64_Re0", make_params("Stokes-C", 2, 64, 8, 2, 2, Partitioner="Skew Cartesian")),
             ("cavity3d_16_Re0", make_params("Stokes-C", 3, 16, 4, 2, 2, Partitioner="Skew Cartesian")),
             ("cavity3d_16_Re0", make_params("Stokes-C", 3, 16, 8, 1))]
    for name, p in cases:
        A, _, _ = load_fixture(name)
        P = hb.Preconditioner(sp.csr_matrix(A), _dictify(p), pattern_only=True)
        P.Initialize()
        st = P.Stats()
        assert st["interior_couplings"] == 0, name
        assert 0 < st["sum_nsd_nb"] <= st["sum_nsd_sq"]
    # a stencil wider than the one-layer separators can decouple (A^2 pattern) must be reported
    A, _, _ = load_fixture("cavity2d_32_Re0")
    A2 = sp.csr_matrix(abs(A) @ abs(A))
    P = hb.Preconditioner(A2, _dictify(make_params("Stokes-C", 2, 32, 4, 1)), pattern_only=True)
    P.Initialize()
    assert P.Stats()["interior_couplings"] > 0


def _compare_all_levels(eqn, dim, nx, ny, nz, sx, levels, cx, part, extra, problem=None):
    import scipy.sparse as sps
    p = make_params(eqn, dim, nx, sx, levels, cx, ny=ny, nz=nz, Partitioner=part, **extra)
    for k, v in (problem or {}).items():
        p.sublist("Problem").set(k, v)
    A = hb.galeri.create_matrix(eqn, dim, nx, ny, nz)
    tv = hb.galeri.create_testvector(A)
    lvl = ohymls.Preconditioner(A, p.copy(), tv)
    try:
        lvl.initialize()
        chain = [lvl]
        for l in range(levels - 1):
            s_ = chain[-1].schur_prec
            nv = len(s_.vsum_gids)
            nxt = ohymls.Preconditioner(sps.identity(nv, format="csr"), p.copy(), np.ones(nv), l + 1, s_.next_hid,
                                        gids=s_.vsum_gids)
            nxt.initialize()
            chain.append(nxt)
    except Exception:
        chain = None                                   # the oracle rejects the configuration ...
    P = hb.Preconditioner(A, _dictify(p), tv, pattern_only=True)
    if chain is None:
        with pytest.raises(hb.HymlsError):             # ... and so must the library (same message family)
            P.Initialize()
        return "rejected"
    P.Initialize()
    for l, lv in enumerate(chain):
        hid = lv.hid
        assert P.NumMySubdomains(l) == hid.num_subdomains()
        for sd in range(hid.num_subdomains()):
            assert np.array_equal(P.GetInteriorGroup(sd, l), np.asarray(hid.interior[sd], dtype=np.int64))
            got = P.GetSeparatorGroups(sd, l)
            assert len(got) == len(hid.groups[sd])
            for (t, g), (to, go) in zip(got, hid.groups[sd]):
                assert t == to and np.array_equal(g, np.asarray(go, dtype=np.int64))
        assert np.array_equal(P.GetMap(api.MAP_SEPARATOR, l), hid.separator_map())
        assert np.array_equal(P.GetMap(api.MAP_VSUM, l), lv.schur_prec.vsum_gids)
    return "equal"


def test_random_grids_and_partitioners_match_oracle():
    """Seeded sweep over non-cubic and ragged grids, both partitioners, 1-2 levels: bit-exact maps on every level,
    and configurations the reference refuses ('not a multiple of the subdomain size', 'subdomain of size 1')
    are refused by both sides."""
    rng = np.random.default_rng(2026)
    outcomes = {"equal": 0, "rejected": 0}
    for _ in range(24):
        dim = int(rng.choice([2, 3]))
        eqn = str(rng.choice(["Laplace", "Stokes-C"]))
        part = str(rng.choice(["Cartesian", "Skew Cartesian"]))
        sx = int(rng.choice([2, 4] if dim == 3 else [2, 4, 8]))
        mult = lambda: int(rng.integers(2, 5 if dim == 3 else 7))   # noqa: E731
        nx, ny = sx * mult(), sx * mult()
        nz = sx * mult() if dim == 3 else 1
        if part == "Cartesian" and rng.random() < 0.4:
            nx += int(rng.integers(1, sx))                          # ragged last subdomain
        levels = int(rng.choice([1, 2]))
        extra = {"Eliminate_Tube_Pressures_With_Velocities": True} \
            if (eqn == "Stokes-C" and dim == 3 and part == "Cartesian") else {}
        outcomes[_compare_all_levels(eqn, dim, nx, ny, nz, sx, levels, 2 if levels > 1 else None, part, extra)] += 1
    assert outcomes["equal"] >= 12 and outcomes["rejected"] >= 1, outcomes


def test_random_pid_maps_match_oracle():
    """BasePartitioner::CreatePIDMap (src/HYMLS_BasePartitioner.cpp:361-586) for random grid sizes, rank counts
    (incl. counts that leave ranks idle) and both partitioners: the subdomain -> rank map is bit-exact."""
    from oracle.skew import SkewCartesianPartitioner
    rng = np.random.default_rng(5)
    checked = 0
    for _ in range(40):
        dim = int(rng.choice([2, 3]))
        sx = int(rng.choice([2, 4, 8] if dim == 2 else [2, 4]))
        nx = sx * int(rng.integers(2, 9 if dim == 2 else 6))
        nprocs = int(rng.choice([1, 2, 3, 4, 6, 8, 16]))
        part = str(rng.choice(["Cartesian", "Skew Cartesian"]))
        p = make_params("Stokes-C", dim, nx, sx, 1, Partitioner=part)
        cls = CartesianPartitioner if part == "Cartesian" else SkewCartesianPartitioner
        try:
            ref = np.asarray(cls(p.copy(), 0, nprocs, 0).partition().pid_map, dtype=np.int32)
        except Exception:
            with pytest.raises(hb.HymlsError):
                hb.pid_map(_dictify(p), nprocs)
            continue
        assert np.array_equal(hb.pid_map(_dictify(p), nprocs), ref), (dim, nx, sx, nprocs, part)
        checked += 1
    assert checked >= 30


@pytest.mark.parametrize("eqn,dim,n,sx,levels,cx,part,extra,problem", [
    # periodic directions (the wrap of the group nodes and of the skew subdomain numbering)
    ("Laplace", 2, (16, 16, 1), 4, 2, 2, "Cartesian", {}, {"x-periodic": True}),
    ("Laplace", 3, (8, 8, 8), 4, 1, None, "Cartesian", {}, {"x-periodic": True, "y-periodic": True, "z-periodic": True}),
    ("Stokes-C", 2, (16, 24, 1), 4, 1, None, "Cartesian", {}, {"y-periodic": True}),
    ("Stokes-C", 2, (16, 16, 1), 4, 2, 2, "Skew Cartesian", {}, {"x-periodic": True}),
    ("Stokes-C", 2, (24, 16, 1), 4, 1, None, "Skew Cartesian", {}, {"x-periodic": True, "y-periodic": True}),
    ("Stokes-C", 3, (8, 8, 8), 4, 1, None, "Skew Cartesian", {}, {"z-periodic": True}),
    ("Stokes-C", 3, (8, 8, 12), 4, 1, None, "Skew Cartesian", {}, {"x-periodic": True, "y-periodic": True, "z-periodic": True}),
    # retained nodes per separator (rx > 1), two retained pressures, unlinked velocities
    ("Laplace", 2, (32, 32, 1), 8, 2, 2, "Cartesian", {"Retain_Nodes": 2}, {}),
    ("Stokes-C", 2, (24, 24, 1), 6, 1, None, "Cartesian", {"Retain_Nodes": 3}, {}),
    ("Stokes-C", 2, (32, 32, 1), 8, 1, None, "Skew Cartesian", {"Retain_Nodes": 2}, {}),
    ("Stokes-C", 3, (8, 8, 8), 4, 1, None, "Skew Cartesian", {"Retain_Nodes": 2, "Eliminate_Velocities_Together": False}, {}),
    ("Stokes-C", 2, (16, 16, 1), 4, 1, None, "Cartesian", {}, {"Retained Pressure Nodes": 2}),
    ("Stokes-C", 2, (16, 16, 1), 4, 1, None, "Skew Cartesian", {}, {"Retained Pressure Nodes": 2}),
    ("Stokes-C", 3, (8, 8, 8), 4, 1, None, "Skew Cartesian", {}, {"Retained Pressure Nodes": 3}),
    ("Stokes-C", 2, (16, 16, 1), 4, 1, None, "Cartesian", {"Eliminate_Velocities_Together": False,
                                                        "Eliminate_Retained_Nodes_Together": False}, {}),
    # subdomain sizes that are not powers of two; coarsening factors with integer roots (4 -> 2, 9 -> 3)
    ("Stokes-C", 2, (36, 36, 1), 6, 2, 3, "Skew Cartesian", {}, {}),
    ("Stokes-C", 3, (12, 12, 12), 6, 1, None, "Skew Cartesian", {}, {}),
    ("Stokes-C", 3, (20, 20, 10), 10, 1, None, "Skew Cartesian", {}, {}),
    ("Laplace", 2, (36, 36, 1), 2, 2, 9, "Cartesian", {}, {}),
    ("Laplace", 3, (16, 16, 16), 2, 2, 4, "Skew Cartesian", {}, {}),
    ("Laplace", 2, (27, 18, 1), 3, 2, 3, "Cartesian", {}, {}),
])
def test_partitioner_variants_match_oracle(eqn, dim, n, sx, levels, cx, part, extra, problem):
    """Parameter variants outside the default sweeps (periodicity, Retain Nodes, several retained pressures,
    unlinked groups, sx = 6 / 10, coarsening 3 / 4 / 9): the geometric formulation of partitioner.cpp gives
    bit-exact maps against the restated reference algorithm on every level."""
    assert _compare_all_levels(eqn, dim, n[0], n[1], n[2], sx, levels, cx, part, extra, problem) == "equal"


@pytest.mark.parametrize("part", ["Cartesian", "Skew Cartesian"])
@pytest.mark.parametrize("cx,nprocs", [(2, 5), (4, 8), (4, 3), (9, 4), (3, 27), (2, 64), (8, 7)])
def test_pid_map_coarsening_factors_match_oracle(part, cx, nprocs):
    from oracle.skew import SkewCartesianPartitioner
    for dim, nx, sx in [(2, 72, 2), (3, 16, 2), (2, 64, 8), (3, 24, 4)]:
        p = make_params("Stokes-C", dim, nx, sx, 1, cx, Partitioner=part)
        cls = CartesianPartitioner if part == "Cartesian" else SkewCartesianPartitioner
        try:
            ref = np.asarray(cls(p.copy(), 0, nprocs, 0).partition().pid_map, dtype=np.int32)
        except Exception:
            with pytest.raises(hb.HymlsError):     # configurations the reference refuses are refused here too
                hb.pid_map(_dictify(p), nprocs)
            continue
        assert np.array_equal(hb.pid_map(_dictify(p), nprocs), ref), (dim, nx, sx, cx, nprocs, part)


def test_initialize_is_independent_of_the_host_thread_count(monkeypatch):
    """The symbolic phase is threaded (per-subdomain group lists, orderings, merge of the hierarchical map, reduced
    Schur pattern): every map and statistic must be identical for 1 and 7 host threads."""
    import scipy.sparse as sp
    from tests.conftest import load_fixture
    A, _, _ = load_fixture("cavity3d_16_Re0")
    p = make_params("Stokes-C", 3, 16, 4, 2, 2, Partitioner="Skew Cartesian")
    seen = []
    for threads in ("1", "7"):
        monkeypatch.setenv("HYMLS_B200_HOST_THREADS", threads)
        P = hb.Preconditioner(sp.csr_matrix(A), _dictify(p), pattern_only=True)
        P.Initialize()
        st = P.Stats()
        maps = [P.GetMap(w, l) for l in range(2)
                for w in (api.MAP_OVERLAPPING, api.MAP_INTERIOR, api.MAP_SEPARATOR, api.MAP_VSUM)]
        seen.append((maps, {k: st[k] for k in ("num_interior", "num_separator", "num_vsum", "num_blocks", "sum_nsd_sq",
                                               "sum_nsd_nb", "interior_couplings")}))
    for a, b in zip(seen[0][0], seen[1][0]):
        assert np.array_equal(a, b)
    assert seen[0][1] == seen[1][1]


def test_header_is_plain_c_and_the_ctypes_mirror_matches_it(tmp_path):
    """The boundary is a C ABI: include/hymls_b200.h compiles as C99, and the ctypes mirrors of its structs
    (hymls_b200/api.py) have the size and the field offsets the C compiler gives them."""
    import subprocess
    root = os.path.dirname(api._HERE)
    src = tmp_path / "layout.c"
    fields = [k for k, _ in api._Stats._fields_]
    info = [k for k, _ in api._SolveInfo._fields_]
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "hymls_b200.h"\nint main(void) {\n'
        '  printf("%zu\\n", sizeof(hymls_b200_stats));\n'
        + "".join('  printf("%%zu\\n", offsetof(hymls_b200_stats, %s));\n' % f for f in fields)
        + '  printf("%zu\\n", sizeof(hymls_b200_solve_info));\n'
        + "".join('  printf("%%zu\\n", offsetof(hymls_b200_solve_info, %s));\n' % f for f in info)
        + "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror",
                           "-I" + os.path.join(root, "include"), str(src), "-o", str(exe)])
    out = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(api._Stats)] + [getattr(api._Stats, f).offset for f in fields] + \
           [ctypes.sizeof(api._SolveInfo)] + [getattr(api._SolveInfo, f).offset for f in info]
    assert out == want
