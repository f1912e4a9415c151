#!/usr/bin/env python
"""Headline benchmark: HYMLS preconditioner hot path on a synthetic 3D lid-driven-cavity (Stokes-C)
Jacobian, BASELINE.json's metric "ApplyInverse/s + HBM GB/s; GMRES solve time".

    python bench.py --gpus N --steps K --warmup W            # this implementation (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the restated reference (oracle)

A "step" is one Preconditioner::ApplyInverse on one right-hand side.  `value` is ApplyInverse/s with
the vectors resident in HBM; `e2e` is the same call through the C ABI with pinned HOST buffers
(H2D of b and D2H of x inside the timed region).  One JSON line is printed by rank 0.

Multi-GPU (torchrun, one rank per GPU): ONE problem, its level-0 subdomains sharded over the ranks by the
reference's subdomain->rank map (CreatePIDMap); partial separator products and interior results are
summed with NCCL all-reduces inside ApplyInverse ("scaling": "strong").  Timing is the max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="hymls_b200", choices=["hymls_b200", "reference"])
    ap.add_argument("--nx", type=int, default=int(os.environ.get("HYMLS_BENCH_NX", 128)))
    ap.add_argument("--sx", type=int, default=int(os.environ.get("HYMLS_BENCH_SX", 8)))
    ap.add_argument("--levels", type=int, default=2)
    ap.add_argument("--cx", type=int, default=4, help="coarsening factor between levels")
    ap.add_argument("--partitioner", default="Skew Cartesian", choices=["Skew Cartesian", "Cartesian"],
                    help="'Skew Cartesian' is what the reference uses for 3D Stokes; 'Cartesian' needs the "
                         "documented tube-pressure extension (DESIGN.md, Deviations)")
    ap.add_argument("--no-solve", action="store_true", help="skip the GMRES solve")
    ap.add_argument("--cpu-sample-nx", type=int, default=16)
    return ap.parse_args()


PARTITIONER = "Skew Cartesian"


def make_params(nx, sx, levels, cx):
    prec = {"Partitioner": PARTITIONER, "Separator Length": sx, "Number of Levels": levels,
            "Coarsening Factor": cx}
    if PARTITIONER == "Cartesian":
        prec["Eliminate Tube Pressures With Velocities"] = True
    return {
        "Problem": {"Equations": "Stokes-C", "Dimension": 3, "nx": nx, "ny": nx, "nz": nx},
        "Preconditioner": prec,
        "Solver": {"Krylov Method": "GMRES", "Initial Vector": "Random", "Left or Right Preconditioning": "Right",
                   "Iterative Solver": {"Maximum Iterations": 600, "Num Blocks": 600, "Maximum Restarts": 0,
                                        "Convergence Tolerance": 1e-8}},
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(alg_bytes):
    """dram__bytes_read.sum + dram__bytes_write.sum of one level-0 k_batched_gemv launch from the committed
    `ncu --set full` capture (profiles/r01_ncu_gemv_traffic.json), if it was taken on this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_gemv_traffic.json")) as f:
            t = json.load(f)
        if abs(t["algorithmic_bytes_per_launch"] - alg_bytes) <= 1e-6 * alg_bytes:
            return t["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def cpu_reference(nx_sample, sx, levels, cx, reps_target_s=8.0):
    """Restated reference (oracle/) on the host: Compute + ApplyInverse on a bounded sample grid."""
    from oracle import hymls as oh
    from oracle.params import ParameterList
    import hymls_b200.galeri as galeri

    def to_pl(d):
        pl = ParameterList()
        for k, v in d.items():
            pl[k] = to_pl(v) if isinstance(v, dict) else v
        return pl

    p = make_params(nx_sample, sx, levels, cx)
    A = -galeri.create_matrix("Stokes-C", 3, nx_sample)
    tv = galeri.create_testvector(A)
    t0 = time.time()
    O = oh.Preconditioner(A, to_pl(p), tv)
    O.initialize()
    O.compute()
    t_compute = time.time() - t0
    b = np.random.default_rng(0).uniform(-1, 1, A.shape[0])
    O.apply_inverse(b)
    reps, t0 = 0, time.time()
    while time.time() - t0 < reps_target_s or reps < 3:
        O.apply_inverse(b)
        reps += 1
    t_apply = (time.time() - t0) / reps
    return {"nsd": O.hid.num_subdomains(), "n": A.shape[0], "apply_s": t_apply, "compute_s": t_compute, "reps": reps}


def _cpu_worker(q, nx_sample, sx, levels, cx, secs):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    try:
        q.put(cpu_reference(nx_sample, sx, levels, cx, secs))
    except Exception as e:  # pragma: no cover
        q.put({"error": repr(e)})


def cpu_reference_all_cores(nx_sample, sx, levels, cx, secs, max_procs=64):
    """One process per host core, each running the restated reference on its own brick of the workload
    concurrently (stands in for `mpirun -np <cores>`, every rank owning a brick).  Returns the list of
    per-process results."""
    import multiprocessing as mp
    cores = max(1, min(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count(),
                       max_procs))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(q, nx_sample, sx, levels, cx, secs)) for _ in range(cores)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
    res = [r for r in res if "error" not in r]
    return cores, res


def main():
    global PARTITIONER
    args = parse()
    PARTITIONER = args.partitioner
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    nx, sx = args.nx, args.sx
    cx = args.cx
    npx = nx // sx
    nsd_full = npx ** 3 if PARTITIONER == "Cartesian" else (2 * npx * npx + 2 * npx) * (npx + 1)
    workload = ("synthetic 3D lid-driven cavity (Stokes-C, GaleriExt::Stokes3D a=nx^2 b=1) %d^3, dof 4, %s partitioner, "
                "sx=%d, %d levels, cx=%d" % (nx, PARTITIONER, sx, args.levels, cx))

    if args.impl == "reference":
        # the reference's own CPU path, restated (the real binary needs Trilinos+MPI: not buildable here)
        if rank != 0:
            return
        snx = min(nx, max(args.cpu_sample_nx, 2 * sx))
        cores, res = cpu_reference_all_cores(snx, sx, min(args.levels, 1) if snx // sx < 4 else args.levels, 2,
                                             max(2.0, 0.4 * args.steps))
        r = res[0]
        nsd_s = r["nsd"]
        # every process advances its own brick; the job-level rate is the sum (work is linear in subdomains)
        v = sum(1.0 / (q["apply_s"] * nsd_full / q["nsd"]) for q in res)
        sample = ("oracle (numpy/scipy SuperLU per subdomain): %d concurrent processes (one per host core), each on "
                  "its own %d^3 brick = %d of the %d subdomains of the workload; rates summed and scaled by the "
                  "subdomain ratio" % (len(res), snx, nsd_s, nsd_full))
        print(json.dumps({
            "impl": "reference", "metric": "apply_inverse_per_s", "value": v, "unit": "1/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "parallelism": "cpu"},
            "cpu_baseline": {"value": v, "unit": "1/s", "cores": cores, "kind": "port", "sample": sample,
                             "sample_apply_ms": r["apply_s"] * 1e3, "sample_compute_s": r["compute_s"]},
            "e2e": {"value": v, "unit": "1/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import hymls_b200 as hb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    t0 = time.time()
    A = -hb.galeri.create_matrix("Stokes-C", 3, nx)
    tv = hb.galeri.create_testvector(A)
    t_gen = time.time() - t0
    n = A.shape[0]
    P = hb.Preconditioner(A, make_params(nx, sx, args.levels, cx), tv)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(hb.Preconditioner.CommUniqueId()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        P.CommInit(bytes(idt.cpu().numpy().tobytes()), rank, world)
    t0 = time.time()
    P.Initialize()
    t_init = time.time() - t0
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()  # ranks leave the (host-side) Initialize at different times: do not bill that to Compute
    t0 = time.time()
    P.Compute()
    torch.cuda.synchronize()
    t_compute = time.time() - t0
    st = P.Stats()

    rng = np.random.default_rng(42)  # the same vectors on every rank (replicated arguments)
    xex = rng.uniform(-1, 1, n)
    bh = A @ xex
    b = torch.from_numpy(bh).cuda()
    x = torch.empty_like(b)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) ----
    for _ in range(max(args.warmup, 3)):
        P.ApplyInverse(b, x)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = P.Stats()["kernel_launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        P.ApplyInverse(b, x)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = P.Stats()["kernel_launches"] - l0
    # ---- end-to-end through the C ABI with pinned host buffers: every rank passes / receives its own rows
    #      of the (distributed) vectors, like the reference's Epetra_MultiVector on an MPI rank ----
    r0, r1 = P.LocalRows()
    hb_in = torch.empty(r1 - r0, dtype=torch.float64).pin_memory()
    hb_out = torch.empty(r1 - r0, dtype=torch.float64).pin_memory()
    hb_in.copy_(torch.from_numpy(bh[r0:r1]))
    bin_np, bout_np = hb_in.numpy(), hb_out.numpy()
    lib, h = P._lib, P._h
    for _ in range(2):
        lib.hymls_b200_apply_inverse_dist(h, bin_np.ctypes.data, bout_np.ctypes.data, 0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lib.hymls_b200_apply_inverse_dist(h, bin_np.ctypes.data, bout_np.ctypes.data, 0)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    e2e_ok = bool(np.allclose(bout_np, x.cpu().numpy()[r0:r1], rtol=0, atol=1e-9 * float(x.abs().max())))
    clocks = sampler.stop()
    # ---- dominant kernel (batched A11^-1 apply, level 0) via CUDA events inside the library ----
    ms_apply_lib, ms_a11 = P.TimeApply(max(5, min(args.steps, 20)))
    st2 = P.Stats()
    # ---- GMRES solve (the other half of the metric) ----
    gm = None
    if not args.no_solve:
        S = hb.Solver(P)
        xs = S.ApplyInverse(b, seed=43)
        torch.cuda.synchronize()
        err = float(np.linalg.norm(xs.cpu().numpy() - xex) / np.linalg.norm(bh))
        gm = {"iterations": S.num_iter, "converged": bool(S.info["converged"]),
              "solve_s": S.info["solve_seconds"], "explicit_rel_residual": S.info["explicit_rel_residual"],
              "rel_error": err, "tol": 1e-8, "restart": "none (Num Blocks 600)"}

    tm = torch.tensor([ms, e2e_ms, ms_a11, t_compute], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms, e2e_ms, ms_a11, t_compute = [float(v) for v in tm.cpu()]
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    # one pass over the explicit A11 inverses this rank owns (8 n_sd^2 bytes per subdomain solve)
    alg_bytes = st["bytes_a11_full_pass"]
    achieved = alg_bytes / (ms_a11 * 1e-3) / 1e9 if ms_a11 > 0 else 0.0
    value = args.steps / (ms * 1e-3)
    out = {
        "metric": "apply_inverse_per_s", "value": value, "unit": "1/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "n": n, "nnz": int(A.nnz), "subdomains": int(st["num_subdomains"]),
                   "parallelism": "level-0 subdomains sharded over %d ranks (CreatePIDMap), NCCL all-reduce of "
                                  "separator / interior vectors; all levels sharded, Krylov vectors replicated" % world
                   if world > 1 else "single GPU",
                   "l2": "inputs larger than L2 (A11 inverses %.2f GB per rank: one full pass + a %.2f GB pass over "
                         "their leading rows per step)" % (alg_bytes / 1e9, (st["bytes_a11_level0"] - alg_bytes) / 1e9),
                   "sum_nsd_sq": st["sum_nsd_sq"], "sum_nsd_nb": st["sum_nsd_nb"],
                   "bytes_apply_algorithmic": st["bytes_apply"],
                   "apply_gbs_all_kernels": st["bytes_apply"] / (ms / args.steps * 1e-3) / 1e9,
                   "t_generate_s": t_gen, "t_initialize_s": t_init, "t_compute_s": t_compute,
                   "compute_tflops": st["flops_compute"] / t_compute / 1e12},
        "e2e": {"value": args.steps / (e2e_ms * 1e-3), "unit": "1/s", "h2d_bytes_per_step": 8 * n,
                "d2h_bytes_per_step": 8 * n, "per_rank_bytes_each_way": 8 * (r1 - r0),
                "matches_device_result": e2e_ok,
                "call": "hymls_b200_apply_inverse_dist (pinned host rows of this rank in / out)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_batched_gemv (A11^-1 apply, level 0, per rank: the full pass "
                                               "of the second subdomain solve)", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(alg_bytes),
                     "bytes_per_launch": alg_bytes, "ms_per_launch": ms_a11, "peak_source": peak_src,
                     "leading_rows_pass": {"bytes_per_launch": st2["bytes_a11_level0"] - alg_bytes,
                                           "ms_per_launch": st2["ms_a11_lead"],
                                           "achieved": ((st2["bytes_a11_level0"] - alg_bytes) /
                                                        (st2["ms_a11_lead"] * 1e-3) / 1e9)
                                           if st2["ms_a11_lead"] > 0 else 0.0}},
        "gmres": gm,
    }
    # CPU baseline on a bounded sample of the same workload (rank 0, N=1 only)
    if world == 1:
        snx = min(nx, max(args.cpu_sample_nx, 2 * sx))
        cores, res = cpu_reference_all_cores(snx, sx, min(args.levels, 1) if snx // sx < 4 else args.levels, 2, 6.0)
        r = res[0]
        v = sum(1.0 / (q["apply_s"] * nsd_full / q["nsd"]) for q in res)
        out["cpu_baseline"] = {
            "value": v, "unit": "1/s", "cores": cores, "kind": "port",
            "sample": "oracle (numpy + scipy SuperLU per subdomain): %d concurrent processes (one per host core), "
                      "each on its own %d^3 brick = %d of %d subdomains; rates summed and scaled by the subdomain "
                      "ratio" % (len(res), snx, r["nsd"], nsd_full),
            "sample_apply_ms": r["apply_s"] * 1e3, "sample_compute_s": r["compute_s"]}
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
