"""GPU, needs >= 2 devices (skipped otherwise): sampling sharded over two ranks (NCCL) returns, on
every rank, exactly the structures of the single-GPU run (noise keyed by global sample id)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, natoms, q, precision):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from chemeleon_b200.config import SamplerConfig
    from chemeleon_b200.dist import sample_sharded
    from chemeleon_b200.sampler import ChemeleonB200
    from chemeleon_b200.weights import random_init_state_dict

    sd = random_init_state_dict(SamplerConfig(), seed=2, head_scale=0.01, lattice_identity=True)
    model = ChemeleonB200(sd, device=f"cuda:{rank}", precision=precision)
    g = torch.Generator().manual_seed(0)
    B = len(natoms)
    text, null = torch.randn(B, 512, generator=g), torch.randn(1, 512, generator=g)
    a, x, l = sample_sharded(model, natoms, text, null, seed=5, t_stop=994)
    q.put((rank, a.cpu(), x.cpu(), l.cpu()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_two_rank_sampling_matches_single_gpu(precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from chemeleon_b200.config import SamplerConfig
    from chemeleon_b200.dist import sample_sharded
    from chemeleon_b200.sampler import ChemeleonB200
    from chemeleon_b200.weights import random_init_state_dict

    natoms = [5, 12, 7, 7, 3, 9]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000 + (7 if precision == "tc" else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, natoms, q, precision)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    sd = random_init_state_dict(SamplerConfig(), seed=2, head_scale=0.01, lattice_identity=True)
    model = ChemeleonB200(sd, device="cuda:0", precision=precision)
    g = torch.Generator().manual_seed(0)
    text, null = torch.randn(len(natoms), 512, generator=g), torch.randn(1, 512, generator=g)
    a1, x1, l1 = sample_sharded(model, natoms, text, null, seed=5, t_stop=994)
    for _, a, x, l in res:
        assert torch.equal(a, a1.cpu())
        assert torch.allclose(x, x1.cpu(), atol=1e-5)
        assert torch.allclose(l, l1.cpu(), atol=1e-5, rtol=1e-5)
