"""TEST INFRASTRUCTURE ONLY - CPU restatement (numpy, float64) of the reference's validity filter.

Only tests/, __graft_entry__.smoke() and bench.py may import this module; the product path
(chemeleon_b200/) never does.

Restates
  * chemeleon/scripts/evaluate.py:177-189 `test_valid`: a structure is dropped when
    `max(st.lattice.abc) > 60` or when the smallest POSITIVE entry of `st.distance_matrix` is
    below 0.5 Angstrom;
  * chemeleon/scripts/sample_target_composition.py:57-62: dropped when `max(abc) > 60` or when
    `st.composition.reduced_composition.alphabetical_formula != comp`.

`Structure.distance_matrix` is pymatgen (unpinned in the reference's requirements.txt, not
installed in this image): the periodic minimum-image distance of every pair of sites
(`Lattice.get_all_distances` -> `pbc_shortest_vectors`: LLL-reduce the cell, scan the 27
neighbouring images).  That quantity is a geometric invariant, so it is restated here by brute
force: minimum of |(x_j - x_i + m) L| over all integer images m in a box large enough to contain
every image closer than `search_radius` (default 3 A, far above the 0.5 A threshold).
PARITY UNPINNED for this function: the reference has no test or fixture for it and pymatgen cannot
be imported here; the known-answer cases in tests/test_validity.py are hand-computed.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import numpy as np

INVALID_LATTICE, INVALID_DISTANCE, INVALID_COMPOSITION = 1, 2, 4


def reduced_counts(numbers: Sequence[int]) -> np.ndarray:
    """Atoms per atomic number divided by their gcd (Composition.reduced_composition)."""
    z = np.asarray(numbers, dtype=np.int64).copy()
    z[(z > 103) | (z < 0)] = 0                      # schema.py:60-62
    h = np.bincount(z, minlength=104)[:104]
    g = 0
    for c in h:
        g = math.gcd(g, int(c))
    return h // g if g > 0 else h


def min_positive_distance(frac: np.ndarray, cell: np.ndarray, search_radius: float = 3.0) -> float:
    """Smallest positive minimum-image distance between two different sites (inf if none)."""
    frac = np.asarray(frac, dtype=np.float64).reshape(-1, 3)
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    n = len(frac)
    if n < 2:
        return math.inf
    vol = abs(np.linalg.det(cell))
    a, b, c = cell
    spacing = [vol / max(np.linalg.norm(np.cross(b, c)), 1e-300), vol / max(np.linalg.norm(np.cross(c, a)), 1e-300),
               vol / max(np.linalg.norm(np.cross(a, b)), 1e-300)]
    R = [min(12, max(1, int(math.ceil(search_radius / max(h, 1e-12) + 0.5)))) for h in spacing]
    imgs = np.array([[i, j, k] for i in range(-R[0], R[0] + 1) for j in range(-R[1], R[1] + 1)
                     for k in range(-R[2], R[2] + 1)], dtype=np.float64)
    best = math.inf
    for i in range(n):
        d = frac[i + 1:] - frac[i]
        d -= np.rint(d)
        v = (d[:, None, :] + imgs[None, :, :]) @ cell          # [n-i-1, images, 3]
        dist = np.sqrt((v * v).sum(-1)).min(axis=1)             # minimum image per pair
        pos = dist[dist > 0]
        if len(pos):
            best = min(best, float(pos.min()))
    return best


def validity_flags(numbers, frac, cell, target_reduced: Optional[np.ndarray] = None, max_length: float = 60.0,
                   min_distance: float = 0.5) -> Dict[str, float]:
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    max_abc = float(np.sqrt((cell * cell).sum(1)).max())
    dmin = min_positive_distance(frac, cell)
    f = 0
    if not max_abc <= max_length:
        f |= INVALID_LATTICE
    if dmin < min_distance:
        f |= INVALID_DISTANCE
    if target_reduced is not None and not np.array_equal(reduced_counts(numbers), np.asarray(target_reduced)[:104]):
        f |= INVALID_COMPOSITION
    return {"flags": f, "min_dist": dmin, "max_abc": max_abc}
