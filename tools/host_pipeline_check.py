"""GPU check of the pipelined host-buffer ApplyInverse (one GPU, pinned buffers, no torch):
serial copies (HYMLS_B200_HOST_PIPELINE=0) vs overlapped copies on the same handle -- results must be bitwise equal --
and the time per call of both, for a few chunk counts.  Appends one JSON object per grid to the output file as it goes.

    python tools/host_pipeline_check.py gpurun_out/host_pipeline.jsonl 64 128
    python tools/host_pipeline_check.py gpurun_out/host_pipeline.jsonl 128:8,8t,6t,10t
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hymls_b200 as hb  # noqa: E402

T0 = time.time()


def log(msg):
    print("[%.1f s] %s" % (time.time() - T0, msg), flush=True)


def pinned(n):
    rt = C.CDLL("libcudart.so.12")
    p = C.c_void_p()
    rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    rc = rt.cudaHostAlloc(C.byref(p), 8 * n, 0)
    assert rc == 0, rc
    return np.frombuffer((C.c_double * n).from_address(p.value), dtype=np.float64)


def params(nx, sx=8):
    return {"Problem": {"Equations": "Stokes-C", "Dimension": 3, "nx": nx, "ny": nx, "nz": nx},
            "Preconditioner": {"Partitioner": "Skew Cartesian", "Separator Length": sx, "Number of Levels": 2,
                               "Coarsening Factor": 4}}


def timed(lib, h, b, x, reps):
    for _ in range(2):
        rc = lib.hymls_b200_apply_inverse_dist(h, b.ctypes.data, x.ctypes.data, 0)
        assert rc == 0, lib.hymls_b200_last_error()
    t = time.perf_counter()
    for _ in range(reps):
        lib.hymls_b200_apply_inverse_dist(h, b.ctypes.data, x.ctypes.data, 0)
    return (time.perf_counter() - t) / reps * 1e3


def run(nx, out, chunk_list, reps):
    os.environ["HYMLS_B200_HOST_PIPELINE_MIN_ROWS"] = "0"
    A = -hb.galeri.create_matrix("Stokes-C", 3, nx)
    tv = hb.galeri.create_testvector(A)
    n = A.shape[0]
    log("%d^3 generated" % nx)
    b, xs, xp = pinned(n), pinned(n), pinned(n)
    b[:] = np.random.default_rng(0).uniform(-1, 1, n)
    rec = {"nx": nx, "n": int(n), "chunks": {}}
    P = None
    for K in chunk_list:   # "8": equal chunks, "8t": tapered towards both ends
        os.environ["HYMLS_B200_HOST_PIPELINE_TAPER"] = "1" if str(K).endswith("t") else "0"
        os.environ["HYMLS_B200_HOST_PIPELINE_CHUNKS"] = str(K).rstrip("t")
        if P is not None:
            P.__del__()   # one handle at a time: 37.6 GB of inverses at 128^3
        P = hb.Preconditioner(A, params(nx), tv)
        P.Initialize()
        P.Compute()
        lib, h = P._lib, P._h
        log("K=%s initialized + computed" % K)
        os.environ["HYMLS_B200_HOST_PIPELINE"] = "0"
        xs[:] = 0
        ms_serial = timed(lib, h, b, xs, reps)
        os.environ["HYMLS_B200_HOST_PIPELINE"] = "1"
        xp[:] = 0
        calls0 = P.Stats()["num_apply_inverse"]
        ms_piped = timed(lib, h, b, xp, reps)
        st = P.Stats()
        # pageable buffers take the serial path whatever the switch says
        xq, bq = np.zeros(n), np.array(b)
        lib.hymls_b200_apply_inverse_dist(h, bq.ctypes.data, xq.ctypes.data, 0)
        r = {"plan_chunks": int(st["host_pipeline_chunks"]), "state": int(st["host_pipeline_state"]),
             "bitwise_equal": bool(np.array_equal(xs, xp)), "pageable_equal": bool(np.array_equal(xs, xq)),
             "finite": bool(np.isfinite(xp).all()), "calls_counted": int(st["num_apply_inverse"] - calls0),
             "calls_made": reps + 2, "ms_serial": ms_serial, "ms_pipelined": ms_piped}
        rec["chunks"][str(K)] = r
        log("K=%s %s" % (K, json.dumps(r)))
        with open(out, "a") as f:
            f.write(json.dumps({"nx": nx, "K": K, **r}) + "\n")
    return rec


if __name__ == "__main__":
    out = sys.argv[1]
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    grids = [v for v in sys.argv[2:]] or ["64"]     # "128" or "128:8,8t,6t" (grid : chunk variants)
    for g in grids:
        nx = int(g.split(":")[0])
        variants = g.split(":")[1].split(",") if ":" in g else (["8"] if nx < 128 else ["8", "16", "4"])
        run(nx, out, variants, 20)
    log("done")
