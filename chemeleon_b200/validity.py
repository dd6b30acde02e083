"""Validity pre-filter of sampled structures on the device.

Mirrors the filters the reference applies after sampling
(chemeleon/scripts/evaluate.py:177-189 `test_valid`; sample_target_composition.py:57-62):
lattice lengths <= 60 A, smallest positive periodic distance >= 0.5 A, and (optionally) reduced
composition equal to the target's.  The reference does this per structure on the CPU through
pymatgen; here one kernel launch (`cb2_validity_filter`, one block per crystal) classifies the
whole batch from the device state, so only structures that pass need to leave the GPU.
"""
from __future__ import annotations

import ctypes as C
import math
import re
from typing import List, Optional, Sequence

import torch

from . import _lib
from .atoms import SYMBOLS

INVALID_LATTICE, INVALID_DISTANCE, INVALID_COMPOSITION = 1, 2, 4
_TOKEN = re.compile(r"([A-Z][a-z]?)(\d*)")


def parse_formula(formula: str) -> List[int]:
    """'TiO2' / 'Li Mn O4' -> atoms per atomic number [104] (no brackets, integer counts)."""
    counts = [0] * 104
    text = formula.replace(" ", "")
    pos = 0
    for m in _TOKEN.finditer(text):
        if m.start() != pos:
            raise ValueError(f"cannot parse formula {formula!r}")
        sym, num = m.group(1), m.group(2)
        if sym not in SYMBOLS[1:]:
            raise ValueError(f"unknown element {sym!r} in {formula!r}")
        counts[SYMBOLS.index(sym)] += int(num) if num else 1
        pos = m.end()
    if pos != len(text) or sum(counts) == 0:
        raise ValueError(f"cannot parse formula {formula!r}")
    return counts


def reduced_formula_counts(formula: str) -> List[int]:
    """Counts of the reduced formula (`Composition(...).reduced_composition`)."""
    counts = parse_formula(formula)
    g = 0
    for c in counts:
        g = math.gcd(g, c)
    return [c // g for c in counts]


def validity_flags(atom_types: torch.Tensor, frac_coords: torch.Tensor, lattices: torch.Tensor,
                   natoms: Sequence[int], target: Optional[str] = None, max_length: float = 60.0,
                   min_distance: float = 0.5):
    """Classify every crystal of a batch on the device.

    atom_types int64[N], frac_coords f32[N,3], lattices f32[B,3,3] (rows = lattice vectors) on a
    CUDA device.  Returns (flags int32[B], min_dist f32[B], max_abc f32[B]); flags == 0 means the
    structure passes every test of the reference's filters."""
    lib = _lib.load()
    dev = atom_types.device
    if dev.type != "cuda":
        raise _lib.Cb2Error("validity_flags needs CUDA tensors (there is no CPU fallback)")
    B = len(natoms)
    off = torch.zeros(B + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(torch.as_tensor(list(natoms), dtype=torch.int64), 0).to(torch.int32)
    if int(off[-1]) != atom_types.numel():
        raise ValueError("natoms does not add up to the number of atoms")
    off = off.to(dev)
    a = atom_types.contiguous().to(torch.int64)
    x = frac_coords.contiguous().to(torch.float32)
    lat = lattices.contiguous().to(torch.float32).reshape(B, 9)
    tgt = None
    if target is not None:
        tgt = torch.tensor(reduced_formula_counts(target), dtype=torch.int32, device=dev)
    flags = torch.empty(B, dtype=torch.int32, device=dev)
    dmin = torch.empty(B, dtype=torch.float32, device=dev)
    abc = torch.empty(B, dtype=torch.float32, device=dev)
    _lib.check(lib.cb2_validity_filter(a.data_ptr(), x.data_ptr(), lat.data_ptr(), off.data_ptr(), B,
                                       tgt.data_ptr() if tgt is not None else None, C.c_float(max_length),
                                       C.c_float(min_distance), flags.data_ptr(), dmin.data_ptr(), abc.data_ptr(),
                                       torch.cuda.current_stream(dev).cuda_stream), "cb2_validity_filter")
    return flags, dmin, abc
