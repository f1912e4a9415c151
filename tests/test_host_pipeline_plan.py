"""Schedule of the pipelined host-buffer ApplyInverse (HostPipePlan, hymls_b200/csrc/symbolic.hpp): on one GPU the
copy of b overlaps the leading-rows pass over the level-0 inverses and the copy of x the full pass.  The schedule is
host work done in Initialize, so it is checked here without a GPU, from its definition:
  * chunk c = consecutive subdomains; its work items are exactly the slabs of its subdomains;
  * every row a chunk reads from b lies below inRows[c+1] (copied before the chunk starts);
  * no row is copied out before the chunk that writes it has run (row >= outRows[chunk of the row])."""
import os

import numpy as np
import pytest

import hymls_b200 as hb


def _plan(P):
    return {k: P.DebugArray("pipe_" + k).astype(np.int64) for k in ("matstart", "leaditem", "fullitem", "inrows", "outrows")}


@pytest.mark.parametrize("partitioner,nx,sx,chunks,taper", [("Skew Cartesian", 16, 4, 8, 0), ("Cartesian", 16, 4, 5, 0),
                                                            ("Skew Cartesian", 32, 8, 16, 0),
                                                            ("Skew Cartesian", 32, 4, 8, 1)])
def test_schedule_follows_its_definition(partitioner, nx, sx, chunks, taper, monkeypatch):
    monkeypatch.setenv("HYMLS_B200_HOST_PIPELINE_MIN_ROWS", "0")
    monkeypatch.setenv("HYMLS_B200_HOST_PIPELINE_TAPER", str(taper))
    monkeypatch.setenv("HYMLS_B200_HOST_PIPELINE_CHUNKS", str(chunks))
    prec = {"Partitioner": partitioner, "Separator Length": sx, "Number of Levels": 2, "Coarsening Factor": 2}
    if partitioner == "Cartesian":
        prec["Eliminate Tube Pressures With Velocities"] = True
    params = {"Problem": {"Equations": "Stokes-C", "Dimension": 3, "nx": nx, "ny": nx, "nz": nx}, "Preconditioner": prec}
    A = -hb.galeri.create_matrix("Stokes-C", 3, nx)
    P = hb.Preconditioner(A, params, pattern_only=True)
    P.Initialize()
    st = P.Stats()
    K = st["host_pipeline_chunks"]
    assert 2 <= K <= chunks and st["host_pipeline_state"] == 0
    pl = _plan(P)
    n = A.shape[0]
    nsd = P.NumMySubdomains(0)
    for k in pl:
        assert len(pl[k]) == K + 1 and np.all(np.diff(pl[k]) >= 0) and pl[k][0] == 0
    assert pl["matstart"][-1] == nsd and np.all(np.diff(pl["matstart"]) > 0)
    assert pl["inrows"][-1] == n and pl["outrows"][-1] == n
    rows_per_item = 32
    chunk_of = np.searchsorted(pl["matstart"], np.arange(nsd), side="right") - 1
    full = np.zeros(K, dtype=np.int64)
    written_by = np.full(n, -1)
    sizes = np.zeros(K)
    for sd in range(nsd):
        g = P.GetInteriorGroup(sd, 0)
        c = chunk_of[sd]
        full[c] += -(-len(g) // rows_per_item)
        sizes[c] += float(len(g)) ** 2
        assert g.max() < pl["inrows"][c + 1]          # b is there before the chunk starts
        assert g.min() >= pl["outrows"][c]            # x does not leave before the chunk has written it
        written_by[g] = c
    assert np.array_equal(np.diff(pl["fullitem"]), full)
    # the leading-rows list has at most as many items per chunk as the full list
    assert np.all(np.diff(pl["leaditem"]) <= full) and pl["leaditem"][-1] > 0
    # rows copied out after chunk c: everything an interior wrote there is done (separator rows are final earlier)
    for c in range(K):
        w = written_by[pl["outrows"][c]:pl["outrows"][c + 1]]
        assert np.all(w <= c)
    # chunks of comparable work, and a schedule that actually overlaps: half of b is not needed by the first chunk
    if taper:   # shares 1 2 4 8 8 4 2 1: the end chunks are the small ones
        assert sizes[0] < 0.5 * sizes.mean() and sizes[-1] < 0.5 * sizes.mean() and sizes.max() <= 3.5 * sizes.mean()
    else:
        assert sizes.max() <= 2.5 * sizes.mean()
    assert pl["inrows"][1] <= 0.75 * n and pl["outrows"][K - 1] >= 0.25 * n


def test_no_schedule_for_small_or_exact_problems(monkeypatch):
    monkeypatch.delenv("HYMLS_B200_HOST_PIPELINE_MIN_ROWS", raising=False)
    A = -hb.galeri.create_matrix("Stokes-C", 3, 8)
    base = {"Problem": {"Equations": "Stokes-C", "Dimension": 3, "nx": 8, "ny": 8, "nz": 8}}
    P = hb.Preconditioner(A, dict(base, Preconditioner={"Partitioner": "Skew Cartesian", "Separator Length": 4,
                                                        "Number of Levels": 1}), pattern_only=True)
    P.Initialize()
    assert P.Stats()["host_pipeline_chunks"] == 0      # 2048 rows: below the default threshold
    monkeypatch.setenv("HYMLS_B200_HOST_PIPELINE_MIN_ROWS", "0")
    Q = hb.Preconditioner(A, dict(base, Preconditioner={"Partitioner": "Skew Cartesian", "Separator Length": 4,
                                                        "Number of Levels": 0}), pattern_only=True)
    Q.Initialize()
    assert Q.Stats()["host_pipeline_chunks"] == 0      # Number of Levels = 0: dense Schur complement, no schedule
    monkeypatch.setenv("HYMLS_B200_HOST_PIPELINE", "0")
    R = hb.Preconditioner(A, dict(base, Preconditioner={"Partitioner": "Skew Cartesian", "Separator Length": 4,
                                                        "Number of Levels": 1}), pattern_only=True)
    R.Initialize()
    assert R.Stats()["host_pipeline_state"] == -2      # switched off


@pytest.mark.gpu
def test_pipelined_host_apply_reproduces_the_serial_path(monkeypatch):
    """On the GPU: pinned host vectors, the copies overlapped in chunks (forced on for a small problem), against the
    serial copies on the same handle -- bitwise -- and against the pageable-buffer path; the handle's own first-call
    check must have passed (state 1).  tools/host_pipeline_check.py does the same at 64^3 / 128^3 and times it."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "host_pipeline_check", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools",
                                            "host_pipeline_check.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    monkeypatch.setenv("HYMLS_B200_HOST_PIPELINE_MIN_ROWS", "0")
    nx = 16
    params = {"Problem": {"Equations": "Stokes-C", "Dimension": 3, "nx": nx, "ny": nx, "nz": nx},
              "Preconditioner": {"Partitioner": "Skew Cartesian", "Separator Length": 4, "Number of Levels": 2,
                                 "Coarsening Factor": 2}}
    A = -hb.galeri.create_matrix("Stokes-C", 3, nx)
    P = hb.Preconditioner(A, params, hb.galeri.create_testvector(A))
    P.Initialize()
    P.Compute()
    n = A.shape[0]
    b, xs, xp = tool.pinned(n), tool.pinned(n), tool.pinned(n)
    b[:] = np.random.default_rng(5).uniform(-1, 1, n)
    lib, h = P._lib, P._h
    monkeypatch.setenv("HYMLS_B200_HOST_PIPELINE", "0")
    assert lib.hymls_b200_apply_inverse(h, b.ctypes.data, n, xs.ctypes.data, n, 1, 0) == 0
    assert P.Stats()["host_pipeline_state"] == -2
    monkeypatch.setenv("HYMLS_B200_HOST_PIPELINE", "1")
    calls = P.Stats()["num_apply_inverse"]
    for _ in range(3):
        xp[:] = 0
        assert lib.hymls_b200_apply_inverse(h, b.ctypes.data, n, xp.ctypes.data, n, 1, 0) == 0
    st = P.Stats()
    assert st["host_pipeline_chunks"] >= 2 and st["host_pipeline_state"] == 1
    assert st["num_apply_inverse"] == calls + 3          # the self-check of the first call is not counted
    assert np.array_equal(xs, xp)
    assert np.array_equal(P.ApplyInverse(np.array(b)), xp)   # pageable buffers: serial copies
