"""CPU: the C++ oracle with the reference's F-matrix ordering / static pivots against the general ordering with
partial pivoting, plain FP64 and refined, on the BASELINE configurations at their stated sizes (profiles/r02_accuracy.md,
addendum).    python tools/oracle_ordering_compare.py [0] [1] [2]"""
import os, sys, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hymls_b200 as hb
from oracle import cpp_oracle as oc
from oracle.params import ParameterList
def _pl(d):
    pl = ParameterList()
    for k, v in d.items():
        pl[k] = _pl(v) if isinstance(v, dict) else v
    return pl
CONFIGS = [
    ("laplace.xml", "Laplace", 2, 128, {"Separator Length": 4, "Number of Levels": 2},
     {"Krylov Method": "CG", "tol": 1e-10, "blocks": 300, "restarts": 20}),
    ("stokes2D.xml", "Stokes-C", 2, 128, {"Partitioner": "Skew Cartesian", "Separator Length": 4, "Number of Levels": 2},
     {"Krylov Method": "GMRES", "tol": 1e-10, "blocks": 30, "restarts": 20}),
    ("cavity3D.xml @ 64^3", "Stokes-C", 3, 64,
     {"Partitioner": "Skew Cartesian", "Separator Length": 4, "Number of Levels": 2},
     {"Krylov Method": "GMRES", "tol": 1e-8, "blocks": 300, "restarts": 20}),
]
which = sys.argv[1:] or ["0","1","2"]
for idx in which:
    name,eqn,dim,nx,prec,sol = CONFIGS[int(idx)]
    A = hb.galeri.create_matrix(eqn, dim, nx)
    if eqn == "Stokes-C": A = -A
    A = sp.csr_matrix(A); tv = hb.galeri.create_testvector(A); n=A.shape[0]
    params = {"Problem": {"Equations": eqn, "Dimension": dim, "nx": nx, "ny": nx, "nz": nx if dim == 3 else 1},
              "Preconditioner": dict(prec)}
    P = hb.Preconditioner(A, params, tv, pattern_only=True); P.Initialize()
    maps = oc.maps_from_library(P)
    b = A @ np.random.default_rng(42).uniform(-1, 1, n)
    x0 = np.random.default_rng(43).uniform(-1, 1, n)
    res = {}
    for fm in (False, True):
        for ref in (1, 0):
            t=time.time()
            O = oc.Preconditioner(A, _pl(params), tv, maps, refine_steps=ref, fmatrix_ordering=fm)
            O.compute(); tc=time.time()-t
            x = O.apply_inverse(b)
            out = (x,)
            if ref == 1:
                xc, its, conv, hist, _ = O.solve(b, x0=x0, method=sol["Krylov Method"], tol=sol["tol"], max_iters=500, num_blocks=sol["blocks"], max_restarts=sol["restarts"])
                out = (x, its, conv, np.array(hist[:15]))
            res[(fm,ref)] = out
            print(name, "fmatrix" if fm else "general", "refine", ref, "compute %.1f s"%tc, "nnz", O.stats()["nnz_factors"], flush=True)
            del O
    r = lambda a,b_: np.linalg.norm(a-b_)/np.linalg.norm(b_)
    print(name, "refined: fmatrix vs general  %.2e"%r(res[(True,1)][0], res[(False,1)][0]),
          "| plain fmatrix vs refined %.2e"%r(res[(True,0)][0], res[(True,1)][0]),
          "| plain general vs refined %.2e"%r(res[(False,0)][0], res[(False,1)][0]))
    print(name, "iterations", res[(True,1)][1], res[(False,1)][1], "converged", res[(True,1)][2], res[(False,1)][2],
          "history max rel diff %.2e"%np.max(np.abs(res[(True,1)][3]-res[(False,1)][3])/np.abs(res[(False,1)][3])), flush=True)
