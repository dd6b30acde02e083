"""Multi-GPU sampling: shard by sample, no data-path collective, one all-gather at the end.

Crystals never interact (edges are intra-crystal, cspnet.py:320-324; every update
is per node / per crystal), so the path shards trivially.  One process per GPU
(torchrun); samples are assigned by longest-processing-time-first on the cost
model F(n) ~ n^2 + 2.4 n; noise is keyed by the GLOBAL sample id,
so 1/2/4/8-GPU runs give identical per-sample results.  The only collective is
the final all-gather of the finished structures (16 B/atom + 36 B/crystal).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def sample_cost(n: int) -> float:
    return float(n) * float(n) + 2.4 * float(n)


def partition_samples(natoms: Sequence[int], world_size: int) -> List[List[int]]:
    """Deterministic LPT assignment of sample indices to ranks; each rank's list is
    sorted by (n, index) so equal-size crystals are contiguous (tile regularity)."""
    import heapq

    nat = np.asarray(natoms, dtype=np.int64)
    order = np.lexsort((np.arange(len(nat)), -nat))              # (-n, index)
    if world_size == 1:
        return [np.lexsort((np.arange(len(nat)), nat)).tolist()]
    cost = (nat * nat + 2.4 * nat).tolist()                       # sample_cost, vectorised
    heap = [(0.0, r) for r in range(world_size)]                  # least loaded rank first, ties by rank
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order.tolist():
        load, r = heapq.heappop(heap)
        parts[r].append(i)
        heapq.heappush(heap, (load + cost[i], r))
    for r, p in enumerate(parts):
        pa = np.asarray(p, dtype=np.int64)
        parts[r] = pa[np.lexsort((pa, nat[pa]))].tolist() if len(p) else []
    return parts


def node_slices(natoms: Sequence[int], idx: Sequence[int]) -> np.ndarray:
    """Global node indices of the samples `idx`, in that order."""
    nat = np.asarray(natoms, dtype=np.int64)
    off = np.zeros(len(nat) + 1, dtype=np.int64)
    np.cumsum(nat, out=off[1:])
    idx = np.asarray(idx, dtype=np.int64)
    if idx.size == 0:
        return np.zeros(0, dtype=np.int64)
    lens = nat[idx]
    first = np.zeros(len(idx), dtype=np.int64)                    # position of each sample's first node in the output
    np.cumsum(lens[:-1], out=first[1:])
    return np.repeat(off[idx] - first, lens) + np.arange(int(lens.sum()), dtype=np.int64)


def gather_structures(a: torch.Tensor, x: torch.Tensor, l: torch.Tensor, natoms: Sequence[int],
                      parts: List[List[int]], group=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """All-gather the ranks' finished (a [n_r], x [n_r,3], l [b_r,3,3]) and restore the
    caller's global sample order.  Payload per rank is padded to the largest shard."""
    world = dist.get_world_size(group)
    nat = np.asarray(natoms, dtype=np.int64)
    n_nodes = [int(nat[p].sum()) if len(p) else 0 for p in parts]
    n_graphs = [len(p) for p in parts]
    max_n, max_b = max(n_nodes), max(n_graphs)
    dev = x.device
    # one flat fp32 payload per rank: [types | coords | lattices]
    payload = torch.zeros(max_n * 4 + max_b * 9, dtype=torch.float32, device=dev)
    r = dist.get_rank(group)
    nr, br = n_nodes[r], n_graphs[r]
    payload[:nr] = a.to(torch.float32)
    payload[max_n:max_n + nr * 3] = x.reshape(-1)
    payload[max_n * 4:max_n * 4 + br * 9] = l.reshape(-1)
    out = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(out, payload, group=group)
    N, B = int(nat.sum()), len(nat)
    A = torch.zeros(N, dtype=torch.int64, device=dev)
    X = torch.zeros(N, 3, dtype=torch.float32, device=dev)
    L = torch.zeros(B, 3, 3, dtype=torch.float32, device=dev)
    for k in range(world):
        nk, bk = n_nodes[k], n_graphs[k]
        if bk == 0:
            continue
        nodes = torch.from_numpy(node_slices(nat, parts[k])).to(dev)
        gidx = torch.tensor(parts[k], dtype=torch.int64, device=dev)
        A[nodes] = out[k][:nk].round().to(torch.int64)
        X[nodes] = out[k][max_n:max_n + nk * 3].view(nk, 3)
        L[gidx] = out[k][max_n * 4:max_n * 4 + bk * 9].view(bk, 3, 3)
    return A, X, L


class ShardPlan:
    """Which samples (and nodes) of a global ragged batch this rank owns."""

    def __init__(self, natoms: Sequence[int], world: int, rank: int):
        self.natoms = [int(n) for n in natoms]
        self.world, self.rank = int(world), int(rank)
        self.parts = partition_samples(self.natoms, self.world)
        self.mine = self.parts[self.rank]
        self.my_natoms = [self.natoms[i] for i in self.mine]
        self.my_nodes = node_slices(self.natoms, self.mine)
        self.B, self.N = len(self.natoms), sum(self.natoms)

    def cost(self, r: Optional[int] = None) -> float:
        return sum(sample_cost(self.natoms[i]) for i in self.parts[self.rank if r is None else r])


def prepare_sharded_run(model, plan: ShardPlan, text_embeds: Optional[torch.Tensor],
                        null_text_embeds: Optional[torch.Tensor], cond_scale: float = 2.0, step_lr: float = 1e-5,
                        seed: int = 0, t_start: Optional[int] = None):
    """This rank's `SamplerRun` (captured CUDA graph, conditioning, Philox keys = GLOBAL sample ids)
    with its initial state set from the global-order initial noise (sharding invariant)."""
    with torch.cuda.device(model.device):
        l_T, x_T = model.initial_noise(plan.B, plan.N, seed)   # global order on every rank
        nodes = torch.from_numpy(plan.my_nodes).to(x_T.device)
        gi = torch.tensor(plan.mine, dtype=torch.int64, device=x_T.device)
        from .sampler import TextCondition

        if isinstance(text_embeds, TextCondition):   # shared FiLM rows + the row of every (variant, crystal)
            ro = text_embeds.row_of.to(x_T.device)
            te, ne = TextCondition(text_embeds.rows, torch.cat([ro[gi], ro[plan.B + gi]])), None
        else:
            te = text_embeds[gi.to(text_embeds.device)] if text_embeds is not None else None
            ne = null_text_embeds
            if ne is not None and ne.shape[0] == plan.B and plan.B != 1:
                ne = ne[gi.to(ne.device)]
        run = model.make_run(plan.my_natoms, te, ne, cond_scale, step_lr, None, seed, plan.mine)
        run.init_state(l_T[gi], x_T[nodes], t_start)
    return run


def sample_sharded(model, natoms: Sequence[int], text_embeds: Optional[torch.Tensor],
                   null_text_embeds: Optional[torch.Tensor], cond_scale: float = 2.0, step_lr: float = 1e-5,
                   seed: int = 0, t_stop: int = 0, group=None, gather: bool = True):
    """Every rank samples its shard with `model` (a ChemeleonB200 on its own GPU) and,
    if `gather`, all ranks return the full (a, x, l) in the caller's sample order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    plan = ShardPlan(natoms, world, rank)
    with torch.cuda.device(model.device):
        run = prepare_sharded_run(model, plan, text_embeds, null_text_embeds, cond_scale, step_lr, seed)
        for _ in range(model.cfg.timesteps, t_stop, -1):
            run.step()
        a, x, l = run.get_state()
        model.last_flags = run.flags.clone()
        if not gather or world == 1:
            if world == 1:
                return gather_local(a, x, l, plan.natoms, plan.parts)
            return a, x, l
        return gather_structures(a, x, l, plan.natoms, plan.parts, group)


def gather_local(a, x, l, natoms, parts):
    """world_size == 1: undo the (n, index) sort of the single shard."""
    nat = np.asarray(natoms, dtype=np.int64)
    dev = x.device
    nodes = torch.from_numpy(node_slices(nat, parts[0])).to(dev)
    gidx = torch.tensor(parts[0], dtype=torch.int64, device=dev)
    A = torch.zeros_like(a)
    X = torch.zeros_like(x)
    L = torch.zeros_like(l)
    A[nodes], X[nodes], L[gidx] = a, x, l
    return A, X, L
