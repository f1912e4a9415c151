"""Cost of k right-hand sides in one ApplyInverse call relative to one (device-resident vectors, CUDA events).
    python tools/multi_rhs_timing.py [nx]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hymls_b200 as hb  # noqa: E402
import bench  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
A, tv, xex, b, x0 = bench.problem(nx)
P = hb.Preconditioner(A, bench.make_params(nx, 8, 2, 4), tv)
P.Initialize()
P.Compute()
n = A.shape[0]
res = {"nx": nx}
for k in (1, 2, 3, 4, 8):
    B = torch.from_numpy(np.random.default_rng(k).uniform(-1, 1, (k, n))).cuda()
    X = torch.empty_like(B)
    for _ in range(3):
        P.ApplyInverse(B if k > 1 else B[0], X if k > 1 else X[0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        P.ApplyInverse(B if k > 1 else B[0], X if k > 1 else X[0])
    e1.record()
    torch.cuda.synchronize()
    res["ms_%d_rhs" % k] = e0.elapsed_time(e1) / 10
for k in (2, 3, 4, 8):
    res["ratio_%d_vs_1" % k] = res["ms_%d_rhs" % k] / res["ms_1_rhs"]
print("MULTI_RHS " + json.dumps(res))
