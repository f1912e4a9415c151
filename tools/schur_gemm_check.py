"""Dense (two-GEMM) Schur rows on the coarser level against the sparse path, end to end on configs/cavity3D.xml (64^3)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hymls_b200 import driver
xml = open(os.path.join(ROOT, "configs", "cavity3D.xml")).read()
res = {}
for mode in ("0", "1"):
    os.environ["HYMLS_B200_SCHUR_GEMM"] = mode
    out = driver.run(xml, {}, None, verbose=False)
    res[mode] = out
    print("SCHUR_GEMM=%s: its %d converged %s residual %.3e compute %.3f s" % (
        mode, out["iterations"], out["converged"], out["residual"], out["t_compute_s"]), flush=True)
xa, xb = res["0"]["_objects"][3], res["1"]["_objects"][3]
print("x rel diff %.3e" % (np.linalg.norm(xa - xb) / np.linalg.norm(xa)))
