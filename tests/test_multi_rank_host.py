"""world_size-2 (gloo, CPU) test of the host-side sharding logic: every rank derives the subdomains it
owns from the reference's subdomain -> rank map; together they partition the subdomains exactly like
BasePartitioner::CreatePIDMap (oracle) says."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import hymls_b200 as hb
from oracle.partitioner import CartesianPartitioner
from tests.common import make_params


def _dictify(p):
    return {k: (_dictify(v) if isinstance(v, dict) else v) for k, v in p.items()}


def _worker(rank, world, port, nx, sx, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = make_params("Stokes-C", 3, nx, sx, 2, 2, Eliminate_Tube_Pressures_With_Velocities=True)
    A = -hb.galeri.create_matrix("Stokes-C", 3, nx)
    P = hb.Preconditioner(A, _dictify(p), pattern_only=True)
    P.SetRank(rank, world)
    P.Initialize()
    own = torch.zeros(P.NumMySubdomains(0), dtype=torch.int32)
    own[torch.from_numpy(P.OwnedSubdomains(0).astype(np.int64))] = rank + 1
    gathered = [torch.zeros_like(own) for _ in range(world)]
    dist.all_gather(gathered, own)
    owner = torch.stack(gathered).sum(0)          # every subdomain claimed by exactly one rank
    ref = np.asarray(CartesianPartitioner(p.copy(), 0, world, 0).partition().pid_map)
    ok = bool(((torch.stack(gathered) > 0).sum(0) == 1).all()) and np.array_equal(owner.numpy() - 1, ref)
    # deeper levels distribute their subdomains by the same rule (8 level-1 subdomains on 2 ranks -> 4 each); only
    # their separator-side work is replicated
    n1 = torch.tensor([len(P.OwnedSubdomains(1))])
    dist.all_reduce(n1)
    ok = ok and int(n1.item()) == P.NumMySubdomains(1) and 0 < len(P.OwnedSubdomains(1)) < P.NumMySubdomains(1)
    # the index maps do not depend on the number of ranks
    ok = ok and P.GetMap(hb.api.MAP_SEPARATOR, 0).shape[0] > 0
    # rows of the distributed-vector entry point (hymls_b200_owned_rows): every row has exactly one owner, and a
    # rank owns the interior rows of its subdomains
    rows = P.OwnedRows()
    cnt = torch.zeros(A.shape[0], dtype=torch.int32)
    cnt[torch.from_numpy(rows)] += 1
    dist.all_reduce(cnt)
    ok = ok and bool((cnt == 1).all()) and bool(np.all(np.diff(rows) > 0))
    mine = set(rows.tolist())
    ok = ok and all(int(g) in mine for sd in P.OwnedSubdomains(0)[:5] for g in P.GetInteriorGroup(int(sd), 0))
    try:
        P.LocalRows()
        ok = False                 # contiguous row blocks no longer exist with several ranks
    except hb.HymlsError:
        pass
    # the skew partitioner (the bench configuration) shards by the same map
    p2 = make_params("Stokes-C", 3, nx, sx, 2, 2, Partitioner="Skew Cartesian")
    Q = hb.Preconditioner(A, _dictify(p2), pattern_only=True)
    Q.SetRank(rank, world)
    Q.Initialize()
    pm = hb.pid_map(_dictify(p2), world)
    mine = np.sort(Q.OwnedSubdomains(0))
    cnt = torch.tensor([len(mine)])
    dist.all_reduce(cnt)
    ok = ok and int(cnt.item()) == Q.NumMySubdomains(0) == len(pm) and len(mine) == int((pm == rank).sum())
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_ranks_partition_the_subdomains():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, _free_port_cached(), 16, 4, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1


_PORT = None


def _free_port_cached():
    global _PORT
    if _PORT is None:
        _PORT = _free_port()
    return _PORT
