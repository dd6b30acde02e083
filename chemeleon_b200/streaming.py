"""Per-timestep streaming without stalling the GPU.

The reference yields `List[ase.Atoms]` after every timestep from a blocking device->host copy
(`TrajectoryContainer.get_atoms`, schema.py:57-83, chemeleon.py:467), and its server streams JSON of
those objects (app/server.py:49-52).  Here a tiny kernel (`cb2_pack_frame`) packs the state of a
timestep into one compact frame on the device, an async copy on a SIDE stream moves it into a pinned
ring buffer, and the sampling stream keeps replaying its CUDA graph: the host only waits for frames
that are `depth` timesteps old.  Frame layout: include/chemeleon_b200.h (`cb2_pack_frame`).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np
import torch

from . import _lib
from .atoms import state_to_atoms

FRAME_HEADER = 16


class Frame:
    """One timestep of a batch in wire format (a numpy uint8 buffer, `to_bytes()` to send it)."""

    def __init__(self, buf: np.ndarray):
        self.buf = buf
        hdr = buf[:FRAME_HEADER].view(np.int32)
        self.t, self.n_nodes, self.n_graphs = int(hdr[0]), int(hdr[1]), int(hdr[2])
        npad = (self.n_nodes + 3) & ~3
        o = FRAME_HEADER
        self.types = buf[o:o + self.n_nodes]
        o += npad
        self.frac_coords = buf[o:o + 12 * self.n_nodes].view(np.float32).reshape(-1, 3)
        o += 12 * self.n_nodes
        self.lattices = buf[o:o + 36 * self.n_graphs].view(np.float32).reshape(-1, 3, 3)

    def to_bytes(self) -> bytes:
        return self.buf.tobytes()

    @classmethod
    def from_bytes(cls, data: bytes) -> "Frame":
        return cls(np.frombuffer(data, dtype=np.uint8).copy())

    def to_atoms(self, natoms: Sequence[int]) -> List:
        return state_to_atoms(self.types.astype(np.int64), self.frac_coords, self.lattices.reshape(-1, 9), natoms)


class FrameStreamer:
    """Ring of `depth` (device frame, pinned host frame) pairs fed from a `SamplerRun`."""

    def __init__(self, run, depth: int = 4):
        self.run = run
        self.lib = run.eng.lib
        dev = run.eng.device
        self.depth = max(1, int(depth))
        self.nbytes = int(self.lib.cb2_frame_bytes(run.N, run.B))
        self.dev_frames = [torch.empty(self.nbytes, dtype=torch.uint8, device=dev) for _ in range(self.depth)]
        self.host_frames = [torch.empty(self.nbytes, dtype=torch.uint8).pin_memory() for _ in range(self.depth)]
        self.packed = [torch.cuda.Event() for _ in range(self.depth)]
        self.copied = [torch.cuda.Event() for _ in range(self.depth)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.pushed = 0
        self.popped = 0

    def push(self) -> None:
        """Enqueue: pack the current state on the sampling stream, copy it out on the side stream."""
        if self.pushed - self.popped >= self.depth:
            raise RuntimeError("FrameStreamer: ring full (pop before pushing)")
        k = self.pushed % self.depth
        main = torch.cuda.current_stream(self.run.eng.device)
        # the slot's previous copy has been consumed by pop(); the device frame is free once that copy ran
        main.wait_event(self.copied[k]) if self.pushed >= self.depth else None
        _lib.check(self.lib.cb2_pack_frame(self.run.topo.byref(), C.byref(self.run.state), self.dev_frames[k].data_ptr(),
                                           self.nbytes, main.cuda_stream), "cb2_pack_frame")
        self.packed[k].record(main)
        self.copy_stream.wait_event(self.packed[k])
        with torch.cuda.stream(self.copy_stream):
            self.host_frames[k].copy_(self.dev_frames[k], non_blocking=True)
            self.copied[k].record(self.copy_stream)
        self.pushed += 1

    def pop(self) -> Frame:
        """Oldest frame not yet handed out (waits for ITS copy only, not for the sampling stream)."""
        if self.popped >= self.pushed:
            raise RuntimeError("FrameStreamer: nothing to pop")
        k = self.popped % self.depth
        self.copied[k].synchronize()
        self.popped += 1
        return Frame(self.host_frames[k].numpy().copy())

    @property
    def pending(self) -> int:
        return self.pushed - self.popped
