"""Sharded (NCCL, owner-computes halo path) vs single-GPU ApplyInverse / solve on the same problem.
Run under torchrun; prints one JSON line on rank 0 (also used by `bench.py --check`).

    torchrun --nproc-per-node 2 tools/mgpu_check.py [nx sx levels cx partitioner]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hymls_b200 as hb  # noqa: E402


def run_check(nx, sx, levels, cx, partitioner, rank, world, border=False, schur_gemm=None):
    prec = {"Partitioner": partitioner, "Separator Length": sx, "Number of Levels": levels, "Coarsening Factor": cx}
    if partitioner == "Cartesian":
        prec["Eliminate Tube Pressures With Velocities"] = True
    params = {"Problem": {"Equations": "Stokes-C", "Dimension": 3, "nx": nx, "ny": nx, "nz": nx},
              "Preconditioner": prec,
              "Solver": {"Krylov Method": "GMRES", "Initial Vector": "Zero",
                         "Iterative Solver": {"Maximum Iterations": 400, "Convergence Tolerance": 1e-8}}}
    if schur_gemm is not None:
        os.environ["HYMLS_B200_SCHUR_GEMM"] = str(schur_gemm)
    A = -hb.galeri.create_matrix("Stokes-C", 3, nx)
    tv = hb.galeri.create_testvector(A)
    n = A.shape[0]
    def fresh_comm_id():   # an NCCL unique id serves ONE communicator
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(hb.Preconditioner.CommUniqueId()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    P = hb.Preconditioner(A, params, tv)
    P.CommInit(fresh_comm_id(), rank, world)
    P.Initialize()
    P.Compute()
    os.environ.pop("HYMLS_B200_SCHUR_GEMM", None)
    Q = hb.Preconditioner(A, params, tv)   # single-GPU replica of the same problem for comparison
    Q.Initialize()
    Q.Compute()
    rng = np.random.default_rng(0)
    b = rng.uniform(-1, 1, n)
    xq = Q.ApplyInverse(b)
    xs = P.ApplyInverse(b)                       # replicated in / out
    rows = P.OwnedRows()
    xd = P.ApplyInverseDist(b[rows])             # distributed in / out (host buffers)
    xdd = P.ApplyInverseDist(torch.from_numpy(b[rows]).cuda()).cpu().numpy()
    rel = lambda a, r: float(np.linalg.norm(a - r) / np.linalg.norm(r))   # noqa: E731
    res = {"apply_rel_diff": rel(xs, xq), "dist_rel_diff": float(np.linalg.norm(xd - xq[rows]) / np.linalg.norm(xq)),
           "dist_host_eq_device": bool(np.array_equal(xd, xdd)), "owned_rows": int(len(rows)), "n": int(n)}
    # every row has exactly one owner
    cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
    cnt[torch.from_numpy(rows).cuda()] += 1
    dist.all_reduce(cnt)
    res["rows_partitioned"] = bool((cnt == 1).all().item())
    rhs = A @ rng.uniform(-1, 1, n)
    S = hb.Solver(P)
    x = S.ApplyInverse(rhs)
    T = hb.Solver(Q)
    y = T.ApplyInverse(rhs)
    k = min(len(S.history), len(T.history), 15)
    res.update({"iterations_sharded": int(S.num_iter), "iterations_single": int(T.num_iter),
                "history_first15_max_rel_diff": float(np.max(np.abs(S.history[:k] - T.history[:k]) / T.history[:k])),
                "solution_rel_diff": rel(x, y), "converged": bool(S.info["converged"]),
                "explicit_rel_residual": float(S.info["explicit_rel_residual"])})
    # distributed caller: every rank passes an interleaved set of rows (a map unrelated to the owner map), vectors
    # live on that map (what the Trilinos adapter does with an arbitrary Epetra row map)
    import scipy.sparse as sp
    mine = np.arange(rank, n, world, dtype=np.int64)[::-1].copy()      # descending: not even sorted
    Ac = sp.csr_matrix(A)
    R = hb.Preconditioner(None, params)
    R.CommInit(fresh_comm_id(), rank, world)
    R.SetMatrixDist(n, mine, Ac[mine, :])
    R.SetTestVectorDist(mine, tv[mine])
    R.Initialize()
    R.SetRowMap(mine)
    R.Compute()
    xm = R.ApplyInverseMap(b[mine])
    res["map_rel_diff"] = float(np.linalg.norm(xm - xq[mine]) / np.linalg.norm(xq))
    R.SetMatrixDist(n, mine, Ac[mine, :] * 2.0)                      # same pattern, new values: only values move
    R.Compute()
    res["map_rescaled_rel_diff"] = float(np.linalg.norm(2.0 * R.ApplyInverseMap(b[mine]) - xq[mine]) / np.linalg.norm(xq))
    del R
    if border:
        # bordered variant (constant-pressure null space) takes the replicated fallback path
        V = np.zeros((n, 1))
        V[3::4, 0] = 1.0
        V /= np.linalg.norm(V)
        for R in (P, Q):
            R.SetBorder(V)
            R.Compute()
        xb, sb = P.ApplyInverseBordered(b, np.zeros(1))
        xr, sr = Q.ApplyInverseBordered(b, np.zeros(1))
        res["bordered_rel_diff"] = rel(xb, xr)
    ta, tb = P.TimeApply(10), Q.TimeApply(10)
    res.update({"apply_ms_sharded": ta[0], "apply_ms_single": tb[0], "world": world})
    return res


def run_check_variants(rank, world):
    """CG (Laplace) and left-preconditioned GMRES on the owner-distributed Krylov path vs the single-GPU solver"""
    out = {}
    for name, eqn, solver in [
            ("cg_laplace", "Laplace", {"Krylov Method": "CG", "Initial Vector": "Zero",
                                       "Iterative Solver": {"Maximum Iterations": 200, "Convergence Tolerance": 1e-10}}),
            ("gmres_left_stokes", "Stokes-C", {"Krylov Method": "GMRES", "Initial Vector": "Random",
                                              "Left or Right Preconditioning": "Left",
                                              "Iterative Solver": {"Maximum Iterations": 300, "Num Blocks": 40,
                                                                   "Convergence Tolerance": 1e-8}})]:
        nx = 24 if eqn == "Laplace" else 16
        params = {"Problem": {"Equations": eqn, "Dimension": 3, "nx": nx, "ny": nx, "nz": nx},
                  "Preconditioner": {"Partitioner": "Skew Cartesian" if eqn != "Laplace" else "Cartesian",
                                     "Separator Length": 4, "Number of Levels": 2, "Coarsening Factor": 2},
                  "Solver": solver}
        A = hb.galeri.create_matrix(eqn, 3, nx)
        if eqn != "Laplace":
            A = -A
        tv = hb.galeri.create_testvector(A)
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(hb.Preconditioner.CommUniqueId()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        P = hb.Preconditioner(A, params, tv)
        P.CommInit(bytes(idt.cpu().numpy().tobytes()), rank, world)
        P.Initialize()
        P.Compute()
        Q = hb.Preconditioner(A, params, tv)
        Q.Initialize()
        Q.Compute()
        b = A @ np.random.default_rng(3).uniform(-1, 1, A.shape[0])
        S, T = hb.Solver(P), hb.Solver(Q)
        x, y = S.ApplyInverse(b), T.ApplyInverse(b)
        out[name] = {"iterations_sharded": int(S.num_iter), "iterations_single": int(T.num_iter),
                     "converged": bool(S.info["converged"]),
                     "solution_rel_diff": float(np.linalg.norm(x - y) / np.linalg.norm(y))}
    return out


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    a = sys.argv[1:]
    nx = int(a[0]) if len(a) > 0 else 16
    sx = int(a[1]) if len(a) > 1 else 4
    levels = int(a[2]) if len(a) > 2 else 2
    cx = int(a[3]) if len(a) > 3 else 2
    part = a[4] if len(a) > 4 else "Skew Cartesian"
    res = run_check(nx, sx, levels, cx, part, rank, world, border="--border" in a,
                    schur_gemm=1 if "--schur-gemm" in a else None)
    if "--variants" in a:
        res["variants"] = run_check_variants(rank, world)
    if rank == 0:
        print("MGPU_CHECK " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
