"""`python -m hymls_b200.driver params.xml` -- the flow of the reference's driver `hymls_main <params.xml>`
(src/main.cpp:48-535) on top of the C ABI: read the Teuchos XML list, create or read the linear system,
null space -> border, Initialize / Compute, right-hand side from a random exact solution, Krylov solve,
residual and error norms.  Works with the reference's own XML files (testSuite/*.xml,
testSuite/integration_tests/*.xml) and with configs/*.xml.  Under torchrun every rank runs it (one GPU each).

Deviations: random vectors come from numpy's PCG64 (seed 42 exact solution; the library's own
generator with seed 43 for a random initial vector) instead of Epetra's LCG; "Data Directory" may be
`fixture:<name>` (tests/golden/<name>.npz) next to a directory with MatrixMarket files (jac/rhs/sol.mtx).
"""
import json
import os
import sys
import time
import xml.etree.ElementTree as ET

import numpy as np
import scipy.sparse as sp

from . import galeri
from .api import Preconditioner, Solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse_parameter_list(text):
    """Teuchos XML ParameterList -> nested dict (bool / int / double / string leaves)."""
    def conv(t, v):
        if t == "bool":
            return v.strip().lower() in ("1", "true")
        if t in ("int", "long long", "unsigned int", "short"):
            return int(v)
        if t in ("double", "float"):
            return float(v)
        return v

    def walk(node):
        d = {}
        for ch in node:
            if ch.tag == "ParameterList":
                d[ch.get("name")] = walk(ch)
            elif ch.tag == "Parameter":
                d[ch.get("name")] = conv(ch.get("type", "string"), ch.get("value"))
        return d

    return walk(ET.fromstring(text))


def create_nullspace(n, kind, problem):
    """MainUtils::create_nullspace (src/HYMLS_MainUtils.cpp:350-441): columns normalised."""
    dim = problem.get("Dimension", 2)
    dof = problem.get("Degrees of Freedom", dim + 1 if problem.get("Equations") in ("Stokes-C", "Stokes-B") else 1)
    pvar = problem.get("Pressure Variable", dim)
    gid = np.arange(n)
    if kind == "Constant":
        V = np.zeros((n, dof))
        V[gid, gid % dof] = 1.0
    elif kind == "Constant P":
        V = np.zeros((n, 1))
        V[gid % dof == pvar, 0] = 1.0
    elif kind == "Checkerboard":
        nx = problem.get("nx", 1)
        ny = problem.get("ny", nx)
        cell = gid // dof
        i, j, k = cell % nx, (cell // nx) % ny, cell // (nx * ny)
        b = 1 if problem.get("Equations") == "Stokes-B" else 0
        v1 = ((i + j + k * b) % 2).astype(np.float64)
        V = np.zeros((n, 2))
        p = gid % dof == pvar
        V[p, 0] = v1[p]
        V[p, 1] = 1.0 - v1[p]
    else:
        raise ValueError("'Null Space'='%s' not implemented" % kind)
    return V / np.linalg.norm(V, axis=0)


def read_system(datadir):
    """(K, rhs or None, sol or None): `fixture:<name>` or a directory with jac.mtx [rhs.mtx, sol.mtx]."""
    if datadir.startswith("fixture:"):
        z = np.load(os.path.join(ROOT, "tests", "golden", datadir[8:] + ".npz"))
        K = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
        return K, (z["rhs"] if "rhs" in z else None), (z["sol"] if "sol" in z else None)
    from scipy.io import mmread
    K = sp.csr_matrix(mmread(os.path.join(datadir, "jac.mtx")))
    vec = {}
    for name in ("rhs", "sol"):
        f = os.path.join(datadir, name + ".mtx")
        vec[name] = np.asarray(mmread(f)).ravel() if os.path.exists(f) else None
    return K, vec["rhs"], vec["sol"]


def _with_diagonal(K, diag_pos, values):
    """K with its EXISTING diagonal entries replaced (Epetra_CrsMatrix::ReplaceDiagonalValues: rows without a stored
    diagonal entry are left alone, so the sparsity pattern -- and the symbolic phase -- stays)"""
    K2 = K.copy()
    rows = np.nonzero(diag_pos >= 0)[0]
    K2.data[diag_pos[rows]] = values[rows]
    return K2


def run(xml_text, overrides=None, comm=None, seed=42, verbose=True):
    """hymls_main (src/main.cpp:48-535): "Number of factorizations" x "Number of solves" (Driver sublist, default
    1 x 1) factorizations / solves; every factorization after a "Diagonal Perturbation" / "Diagonal Shift" of the
    matrix when those are set (:343-360) -- the same-pattern SetMatrix + Compute path NOX takes every Newton step.
    Returns a dict with the numbers hymls_main prints for the LAST solve and the list of all of them in "runs".
    `comm` = (unique_id, rank, nranks) for the sharded (one process per GPU) run."""
    params = parse_parameter_list(xml_text)
    for path, v in (overrides or {}).items():
        d = params
        keys = path.split("/")
        for k in keys[:-1]:
            d = d.setdefault(k, {})
        d[keys[-1]] = v
    driver = params.pop("Driver", {})
    problem = params.setdefault("Problem", {})
    eqn = problem.get("Equations", driver.get("Galeri Label", "not-set"))
    dim = problem.get("Dimension", 2)
    nx = problem.get("nx", 32)
    ny = problem.get("ny", nx)
    nz = problem.get("nz", nx if dim > 2 else 1)
    rhs = sol = None
    t0 = time.time()
    if driver.get("Read Linear System", False):
        K, rhs, sol = read_system(driver.get("Data Directory", "not specified"))
        if not driver.get("RHS Available", rhs is not None):
            rhs = None
        if not driver.get("Exact Solution Available", sol is not None):
            sol = None
    else:
        label = "Laplace" if eqn.startswith("Laplace") else eqn
        K = galeri.create_matrix(label, dim, nx, ny, nz)
        if eqn == "Stokes-C":
            K = -K  # "scale equations by -1" (main.cpp:292-297)
    K = sp.csr_matrix(K)
    n = K.shape[0]
    t_matrix = time.time() - t0
    tv = galeri.create_testvector(K)
    P = Preconditioner(K, params, tv)
    if comm is not None:
        P.CommInit(*comm)
    t0 = time.time()
    P.Initialize()
    t_init = time.time() - t0
    V = None
    kind = driver.get("Null Space Type", "None")
    if kind != "None":
        V = create_nullspace(n, kind, problem)
        P.SetBorder(V)  # solver->SetBorder(nullSpace), main.cpp:363-366
    num_computes = max(1, int(driver.get("Number of factorizations", 1)))
    num_solves = max(1, int(driver.get("Number of solves", 1)))
    perturbation = float(driver.get("Diagonal Perturbation", 0.0))
    diag_shift = float(driver.get("Diagonal Shift", 0.0))
    K0 = K
    perturb = perturbation != 0.0 or diag_shift != 0.0
    if perturb:
        K0.sort_indices()
        diag_pos = np.full(n, -1, dtype=np.int64)   # position of every stored diagonal entry in K.data
        rows = np.repeat(np.arange(n), np.diff(K0.indptr))
        on_diag = np.nonzero(K0.indices == rows)[0]
        diag_pos[rows[on_diag]] = on_diag
        diag0 = K0.diagonal()
        rng_pert = np.random.default_rng(seed + 1000)
    rng = np.random.default_rng(seed)
    S = Solver(P)
    read_rhs = rhs
    runs = []
    show = verbose and (comm is None or comm[1] == 0)
    for f in range(num_computes):
        if perturb:  # "change the matrix values just to see if that works" (main.cpp:343-358)
            K = _with_diagonal(K0, diag_pos, diag0 + diag_shift + perturbation * rng_pert.uniform(-1, 1, n))
            P.SetMatrix(K)          # same pattern: the symbolic phase of Initialize() is kept
        if V is not None and f > 0:
            P.SetBorder(V)          # solver->SetBorder(nullSpace) before every Compute (main.cpp:361-364)
        t0 = time.time()
        P.Compute()
        t_compute = time.time() - t0
        for s_ in range(num_solves):
            if read_rhs is None:
                x_ex = rng.uniform(-1, 1, n)
                if V is not None:
                    x_ex -= V @ (V.T @ x_ex)  # project the null space out of x_ex (main.cpp:401-409)
                rhs = K @ x_ex
            else:
                x_ex = sol
            t0 = time.time()
            x = S.ApplyInverse(rhs, seed=seed + 1)
            t_solve = time.time() - t0
            res = np.linalg.norm(K @ x - rhs) / np.linalg.norm(rhs)
            out = {"equations": eqn, "n": int(n), "nnz": int(K.nnz), "levels": P.NumLevels(),
                   "subdomains": P.NumMySubdomains(0), "iterations": S.num_iter,
                   "converged": bool(S.info["converged"]), "residual": float(res), "t_matrix_s": t_matrix,
                   "t_initialize_s": t_init, "t_compute_s": t_compute, "t_solve_s": t_solve,
                   "t_solve_device_s": S.info["solve_seconds"], "border": 0 if V is None else V.shape[1],
                   "history": [float(h) for h in S.history]}
            if x_ex is not None:
                err = x - x_ex
                if V is not None:
                    err -= V @ (V.T @ err)
                elif eqn == "Stokes-C" and sol is not None:
                    pv = create_nullspace(n, "Constant P", problem)  # integration_tests.cpp:585-604
                    err -= pv @ (pv.T @ err)
                out["error"] = float(np.linalg.norm(err) / np.linalg.norm(rhs))
            runs.append({"factorization": f + 1, "solve": s_ + 1, "iterations": out["iterations"],
                         "converged": out["converged"], "residual": out["residual"], "error": out.get("error"),
                         "t_compute_s": t_compute, "t_solve_device_s": out["t_solve_device_s"]})
            if show:
                if num_computes * num_solves > 1:
                    print("Compute Preconditioner (%d)  Solve (%d)" % (f + 1, s_ + 1))
                print("Residual Norm ||Ax-b||/||b||: %.8e" % out["residual"])
                if "error" in out:
                    print("Error Norm ||x-x_ex||/||b||: %.8e" % out["error"])
                print("iterations: %d  converged: %s  compute %.3f s  solve %.3f s" % (
                    out["iterations"], out["converged"], t_compute, out["t_solve_device_s"]))
    out["runs"] = runs
    out["_objects"] = (K, P, S, x, rhs)
    return out


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        print("usage: python -m hymls_b200.driver params.xml [Sublist/Name=value ...] [--json out.json]")
        return 2
    xml = open(argv[0]).read()
    overrides, js = {}, None
    it = iter(argv[1:])
    for a in it:
        if a == "--json":
            js = next(it)
            continue
        k, v = a.split("=", 1)
        for cast in (int, float):
            try:
                v = cast(v)
                break
            except ValueError:
                pass
        if v in ("true", "false"):
            v = v == "true"
        overrides[k] = v
    comm = None
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(lr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(Preconditioner.CommUniqueId()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = (bytes(idt.cpu().numpy().tobytes()), rank, world)
    out = run(xml, overrides, comm)
    out.pop("_objects")
    if js and (comm is None or comm[1] == 0):
        with open(js, "w") as f:
            json.dump(out, f)
    return 0


if __name__ == "__main__":
    sys.exit(main())
