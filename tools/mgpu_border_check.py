"""Sharded (NCCL) vs single-GPU runs of the bordered configurations through the driver.  Run with torchrun."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hymls_b200 as hb  # noqa: E402
from hymls_b200 import driver  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    def new_comm():  # one NCCL id per communicator
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(hb.Preconditioner.CommUniqueId()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return (bytes(idt.cpu().numpy().tobytes()), rank, world)
    runs = [("cavity.xml", {}),
            ("cavity3D.xml", {"Problem/nx": 16, "Problem/ny": 16, "Problem/nz": 16,
                              "Preconditioner/Separator Length": 4, "Preconditioner/Coarsening Factor": 2,
                              "Preconditioner/Fix Pressure Level": False, "Driver/Null Space Type": "Constant P"})]
    for name, over in runs:
        xml = open(os.path.join(ROOT, "configs", name)).read()
        a = driver.run(xml, over, new_comm(), verbose=False)
        b = driver.run(xml, over, None, verbose=False)
        Ka, Pa, Sa, xa, rhs = a["_objects"]
        Kb, Pb, Sb, xb, _ = b["_objects"]
        n = Ka.shape[0]
        rng = np.random.default_rng(3)
        B = rng.uniform(-1, 1, n); T = rng.uniform(-1, 1, a["border"])
        Xa, Sa_ = Pa.ApplyInverseBordered(B, T)
        Xb, Sb_ = Pb.ApplyInverseBordered(B, T)
        print("rank %d %s: bordered apply sharded vs single rel diff X %.2e S %.2e | its %d vs %d | residual %.2e vs %.2e "
              "| x rel diff %.2e" % (rank, name, np.linalg.norm(Xa - Xb) / np.linalg.norm(Xb),
                                     np.linalg.norm(Sa_ - Sb_) / np.linalg.norm(Sb_), a["iterations"], b["iterations"],
                                     a["residual"], b["residual"], np.linalg.norm(xa - xb) / np.linalg.norm(xb)),
              flush=True)
        assert abs(a["iterations"] - b["iterations"]) <= 1 and a["converged"]
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
