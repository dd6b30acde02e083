"""Ragged-batch topology for the fully connected per-crystal graphs.

Replaces `CSPNet.gen_edges` (cspnet.py:319-324), which builds a dense [N,N]
block-diagonal adjacency and `nonzero()`s it on every forward.  The topology of a
sampling run never changes, so everything here is computed once on the host from
the `natoms` list and kept on the device for the whole run:

* node -> crystal maps and per-crystal node ranges;
* exact path: the edge rows in the reference's order (crystal by crystal, i
  outer, j inner, self loops included) and chunk boundaries of whole segments;
* tensor-core path: tiles of 128 edge rows made of whole (i, all-j) segments of
  equal length n, so that the segmented mean never straddles a tile.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib

EXACT_CHUNK_EDGES = 1 << 16


class BatchTopology:
    def __init__(self, natoms: Sequence[int], n_variants: int, device, exact: bool = True,
                 tensor_core: bool = True, chunk_edges: int = EXACT_CHUNK_EDGES):
        nat = np.asarray(list(natoms), dtype=np.int64)
        if nat.ndim != 1 or (nat < 1).any():
            raise ValueError("natoms must be a list of positive integers")
        self.natoms = nat
        self.B = int(nat.shape[0])
        self.N = int(nat.sum())
        self.V = int(n_variants)
        self.E = int((nat * nat).sum())
        self.max_n = int(nat.max()) if self.B else 0
        if self.N >= 2 ** 31 - 1 or self.V * self.N >= 2 ** 31 - 1:
            raise ValueError("batch too large for int32 node indices")
        dev = torch.device(device)
        self.device = dev
        goff = np.zeros(self.B + 1, dtype=np.int64)
        np.cumsum(nat, out=goff[1:])
        node2graph = np.repeat(np.arange(self.B, dtype=np.int64), nat)
        node_base = goff[:-1][node2graph]
        node_n = nat[node2graph]
        eoff = np.zeros(self.N + 1, dtype=np.int64)
        np.cumsum(node_n, out=eoff[1:])

        def dev_i32(a):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)

        self.graph_off = dev_i32(goff)
        self.node2graph = dev_i32(node2graph)
        self.node_base = dev_i32(node_base)
        self.node_n = dev_i32(node_n)
        self.node_eoff = torch.from_numpy(eoff).to(dev)
        self._host_keep = []
        self.edge_i = self.edge_j = None
        self.n_chunks = 0
        self.chunk_max_edges = 0
        self.host_chunk_node_lo = self.host_chunk_edge_lo = None
        if exact:
            ei = np.repeat(np.arange(self.N, dtype=np.int64), node_n)
            # j = base(i) + (row - eoff(i))
            ej = node_base[ei] + (np.arange(self.E, dtype=np.int64) - eoff[:-1][ei])
            self.edge_i, self.edge_j = dev_i32(ei), dev_i32(ej)
            lo = [0]
            cur = 0
            limit = max(int(chunk_edges), self.max_n)
            for i in range(self.N):
                if cur + node_n[i] > limit:
                    lo.append(i)
                    cur = 0
                cur += int(node_n[i])
            lo.append(self.N)
            node_lo = np.asarray(lo, dtype=np.int32)
            edge_lo = eoff[node_lo].astype(np.int64)
            self.host_chunk_node_lo = np.ascontiguousarray(node_lo)
            self.host_chunk_edge_lo = np.ascontiguousarray(edge_lo)
            self.n_chunks = len(lo) - 1
            self.chunk_max_edges = int(np.diff(edge_lo).max()) if self.n_chunks else 0
        self.n_tiles = 0
        self.tile_row_i = self.tile_row_j = self.tile_seg_n = None
        if tensor_core and self.max_n <= _lib.TILE_ROWS:
            ri, rj, sn = build_tiles(node_n, node_base)
            self.n_tiles = int(sn.shape[0])
            self.tile_row_i, self.tile_row_j, self.tile_seg_n = dev_i32(ri), dev_i32(rj), dev_i32(sn)
        self.struct = self._make_struct()

    def _make_struct(self) -> _lib.Batch:
        b = _lib.Batch()
        b.n_nodes, b.n_graphs, b.n_variants, b.max_n = self.N, self.B, self.V, self.max_n
        b.n_edges = self.E
        b.node2graph = _lib.ptr(self.node2graph)
        b.node_base = _lib.ptr(self.node_base)
        b.node_n = _lib.ptr(self.node_n)
        b.graph_off = _lib.ptr(self.graph_off)
        b.edge_i = _lib.ptr(self.edge_i)
        b.edge_j = _lib.ptr(self.edge_j)
        b.node_eoff = _lib.ptr(self.node_eoff)
        b.n_chunks = self.n_chunks
        if self.host_chunk_node_lo is not None:
            b.host_chunk_node_lo = self.host_chunk_node_lo.ctypes.data
            b.host_chunk_edge_lo = self.host_chunk_edge_lo.ctypes.data
        b.chunk_max_edges = self.chunk_max_edges
        b.n_tiles = self.n_tiles
        b.tile_row_i = _lib.ptr(self.tile_row_i)
        b.tile_row_j = _lib.ptr(self.tile_row_j)
        b.tile_seg_n = _lib.ptr(self.tile_seg_n)
        return b

    def byref(self):
        return C.byref(self.struct)


def build_tiles(node_n: np.ndarray, node_base: np.ndarray, rows: int = _lib.TILE_ROWS):
    """Pack whole segments (node i with all its n neighbours j) of equal n into
    tiles of `rows` edge rows.  Returns (row_i [T*rows], row_j [T*rows], seg_n [T]);
    padding rows have row_i = -1."""
    ri_all, rj_all, sn_all = [], [], []
    for n in np.unique(node_n):
        n = int(n)
        nodes = np.nonzero(node_n == n)[0]
        S = rows // n
        if S < 1:
            raise ValueError(f"crystal with {n} atoms does not fit a {rows}-row tile")
        T = (len(nodes) + S - 1) // S
        seg = np.full(T * S, -1, dtype=np.int64)
        seg[: len(nodes)] = nodes
        seg = seg.reshape(T, S)
        ri = np.full((T, rows), -1, dtype=np.int64)
        rj = np.zeros((T, rows), dtype=np.int64)
        body_i = np.repeat(seg, n, axis=1)                       # [T, S*n]
        jj = np.tile(np.arange(n, dtype=np.int64), S)[None, :]   # [1, S*n]
        base = np.where(body_i >= 0, node_base[np.clip(body_i, 0, None)], 0)
        ri[:, : S * n] = body_i
        rj[:, : S * n] = np.where(body_i >= 0, base + jj, 0)
        ri_all.append(ri.reshape(-1))
        rj_all.append(rj.reshape(-1))
        sn_all.append(np.full(T, n, dtype=np.int64))
    if not ri_all:
        z = np.zeros(0, dtype=np.int64)
        return z, z, z
    return np.concatenate(ri_all), np.concatenate(rj_all), np.concatenate(sn_all)
