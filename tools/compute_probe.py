"""Compute() of the bench workload with phase timers and nvidia-smi clock samples (diagnostic)."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["HYMLS_B200_VERBOSE"] = "1"
import numpy as np, torch
import hymls_b200 as hb
import bench
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 128
bench.PARTITIONER = sys.argv[2] if len(sys.argv) > 2 else "Skew Cartesian"
A = -hb.galeri.create_matrix("Stokes-C", 3, nx)
P = hb.Preconditioner(A, bench.make_params(nx, 8, 2, 4), hb.galeri.create_testvector(A))
P.Initialize()
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu",
                        "--format=csv,noheader", "-lms", "100"], stdout=subprocess.PIPE, text=True)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    P.Compute()
    torch.cuda.synchronize(); print("Compute %d: %.3f s" % (rep, time.time() - t0), flush=True)
smi.terminate()
lines = smi.stdout.read().strip().splitlines()
print("clock samples:", " | ".join(lines[::3]))
