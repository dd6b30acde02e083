"""CPU, world_size 2 over gloo: the multi-GPU path's host logic (partition + the final
all-gather of finished structures restores the caller's order)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, natoms, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from chemeleon_b200.dist import gather_structures, node_slices, partition_samples

    parts = partition_samples(natoms, world)
    mine = parts[rank]
    nodes = node_slices(natoms, mine)
    # synthetic "finished structures": values encode the global node / sample id
    a = torch.from_numpy(nodes % 104)
    x = torch.from_numpy(np.stack([nodes, nodes + 0.25, nodes + 0.5], 1).astype(np.float32))
    l = torch.tensor(mine, dtype=torch.float32).view(-1, 1, 1).expand(-1, 3, 3).contiguous()
    A, X, L = gather_structures(a, x, l, natoms, parts)
    q.put((rank, A.numpy(), X.numpy(), L.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_structures_world2():
    natoms = [5, 40, 7, 7, 12, 33, 4, 20, 20, 9]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, natoms, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    N = sum(natoms)
    nodes = np.arange(N)
    for _, A, X, L in res:
        assert np.array_equal(A, nodes % 104)
        assert np.allclose(X[:, 0], nodes) and np.allclose(X[:, 2], nodes + 0.5)
        assert np.allclose(L[:, 0, 0], np.arange(len(natoms)))
