"""Sharded (NCCL) vs single-GPU ApplyInverse / solve on the same problem.  Run with torchrun."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hymls_b200 as hb  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    sx = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    levels = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    params = {"Problem": {"Equations": "Stokes-C", "Dimension": 3, "nx": nx, "ny": nx, "nz": nx},
              "Preconditioner": {"Separator Length": sx, "Number of Levels": levels, "Coarsening Factor": int(sys.argv[4]) if len(sys.argv) > 4 else 2,
                                 "Eliminate Tube Pressures With Velocities": True},
              "Solver": {"Krylov Method": "GMRES", "Initial Vector": "Zero",
                         "Iterative Solver": {"Maximum Iterations": 300, "Convergence Tolerance": 1e-8}}}
    A = -hb.galeri.create_matrix("Stokes-C", 3, nx)
    tv = hb.galeri.create_testvector(A)
    n = A.shape[0]
    # unique id from rank 0
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(hb.Preconditioner.CommUniqueId()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    uid = bytes(idt.cpu().numpy().tobytes())
    P = hb.Preconditioner(A, params, tv)
    P.CommInit(uid, rank, world)
    P.Initialize()
    own = P.OwnedSubdomains()
    P.Compute()
    Q = hb.Preconditioner(A, params, tv)   # single-GPU replica for comparison
    Q.Initialize(); Q.Compute()
    b = np.random.default_rng(0).uniform(-1, 1, n)
    xs = P.ApplyInverse(b)
    xq = Q.ApplyInverse(b)
    err = np.linalg.norm(xs - xq) / np.linalg.norm(xq)
    # sequence of device-resident applies (what GMRES does)
    rng = np.random.default_rng(5)
    for k in range(4):
        v = torch.from_numpy(rng.uniform(-1, 1, n)).cuda()
        d = (P.ApplyInverse(v) - Q.ApplyInverse(v)).norm().item() / Q.ApplyInverse(v).norm().item()
        print("rank %d device apply %d rel diff %.2e" % (rank, k, d), flush=True)
    v = rng.uniform(-1, 1, n)
    print("rank %d ApplyMatrix diff %.2e" % (rank, np.linalg.norm(P.ApplyMatrix(v) - A @ v) / np.linalg.norm(A @ v)), flush=True)
    from oracle import krylov as ok
    log = []
    def both(z):
        a_ = P.ApplyInverse(z); b_ = Q.ApplyInverse(z)
        d_ = a_ - b_
        log.append((np.linalg.norm(d_) / np.linalg.norm(b_), np.abs(d_).max(), int(np.abs(d_).argmax()) % 4,
                    np.linalg.norm(d_[3::4]) / max(np.linalg.norm(b_[3::4]), 1e-300)))
        return a_
    xo, its_o, conv_o, h_o = ok.gmres(lambda z: A @ z, A @ b, np.zeros(n), both, side="Right", tol=1e-8,
                                      max_iters=40, max_restarts=0)
    if rank == 0:
        for k, l_ in enumerate(log[:40:3]):
            print("  apply %d: rel diff %.2e maxabs %.2e at var %d, pressure rel diff %.2e" % ((3 * k,) + l_), flush=True)
    print("rank %d python GMRES with sharded ApplyInverse: its %d conv %s" % (rank, its_o, conv_o), flush=True)
    S = hb.Solver(P); x = S.ApplyInverse(A @ b)
    T = hb.Solver(Q); y = T.ApplyInverse(A @ b)
    print("rank %d hist sharded" % rank, S.history[::12], "single", T.history[::12], flush=True)
    ta = P.TimeApply(10); tb = Q.TimeApply(10)
    print("rank %d/%d owns %d of %d sds | sharded vs single apply rel diff %.2e | its %d vs %d | x diff %.2e | "
          "apply ms sharded %.3f single %.3f" % (rank, world, len(own), P.NumMySubdomains(0), err, S.num_iter,
                                                  T.num_iter, np.linalg.norm(x - y) / np.linalg.norm(y), ta[0], tb[0]),
          flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
