"""Output boundary: device state -> list of `ase.Atoms`.

Restates `TrajectoryContainer.get_atoms` (chemeleon/modules/schema.py:57-83):
types > 103 -> 0, split per crystal, `Atoms(numbers, cell=lattice, pbc=True)`,
`set_scaled_positions(frac)`, `ase.build.tools.sort(atoms)` (stable sort by
chemical-symbol string).  If `ase` is not installed a minimal stand-in with the
same attributes is returned.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

SYMBOLS = [
    "X", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar",
    "K", "Ca", "Sc", "Ti", "V", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn", "Ga", "Ge", "As", "Se", "Br", "Kr",
    "Rb", "Sr", "Y", "Zr", "Nb", "Mo", "Tc", "Ru", "Rh", "Pd", "Ag", "Cd", "In", "Sn", "Sb", "Te", "I", "Xe",
    "Cs", "Ba", "La", "Ce", "Pr", "Nd", "Pm", "Sm", "Eu", "Gd", "Tb", "Dy", "Ho", "Er", "Tm", "Yb", "Lu",
    "Hf", "Ta", "W", "Re", "Os", "Ir", "Pt", "Au", "Hg", "Tl", "Pb", "Bi", "Po", "At", "Rn", "Fr", "Ra", "Ac",
    "Th", "Pa", "U", "Np", "Pu", "Am", "Cm", "Bk", "Cf", "Es", "Fm", "Md", "No", "Lr",
]

try:  # pragma: no cover - ase is absent in the offline image
    from ase import Atoms as _AseAtoms
    from ase.build.tools import sort as _ase_sort

    HAVE_ASE = True
except Exception:  # pragma: no cover
    HAVE_ASE = False


_PBC = np.array([True, True, True])
_PBC.setflags(write=False)


class AtomsLite:
    """Stand-in for ase.Atoms when ase is not importable."""

    __slots__ = ("numbers", "cell", "_scaled", "pbc")

    def __init__(self, numbers, cell, scaled_positions, pbc=True):
        self.numbers = np.asarray(numbers, dtype=np.int64)
        self.cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
        self._scaled = np.asarray(scaled_positions, dtype=np.float64).reshape(-1, 3)
        self.pbc = np.array([pbc] * 3)

    @classmethod
    def _from_views(cls, numbers, cell, scaled_positions):
        """Batch boundary: the arguments are already int64 [n] / float64 [3,3] / float64 [n,3] slices of the batch
        arrays (no per-structure conversion: 32 768 structures are built on rank 0 after every multi-GPU job)."""
        o = cls.__new__(cls)
        o.numbers, o.cell, o._scaled, o.pbc = numbers, cell, scaled_positions, _PBC
        return o

    def get_scaled_positions(self):
        return self._scaled

    def get_atomic_numbers(self):
        return self.numbers

    def get_chemical_symbols(self):
        return [SYMBOLS[z] for z in self.numbers]

    def get_cell(self):
        return self.cell

    @property
    def positions(self):
        return self._scaled @ self.cell

    def __len__(self):
        return len(self.numbers)

    def __repr__(self):
        return f"AtomsLite(symbols={''.join(self.get_chemical_symbols())}, n={len(self)})"


def symbol_sort_order(numbers: np.ndarray) -> np.ndarray:
    """Index order of `ase.build.tools.sort`: sorted((symbol, index))."""
    deco = sorted((SYMBOLS[int(z)], i) for i, z in enumerate(numbers))
    return np.array([i for _, i in deco], dtype=np.int64)


# rank of every atomic number in the order of its chemical-symbol STRING (what ase's sort compares)
_SYMBOL_RANK = np.empty(len(SYMBOLS), dtype=np.int64)
_SYMBOL_RANK[np.array(sorted(range(len(SYMBOLS)), key=lambda z: SYMBOLS[z]))] = np.arange(len(SYMBOLS))


def batch_symbol_order(a: np.ndarray, natoms: Sequence[int]) -> np.ndarray:
    """`symbol_sort_order` of every crystal of a batch in one vectorised stable sort: global atom
    indices, crystal by crystal, each crystal's atoms ordered by (symbol string, original index)."""
    nat = np.asarray(natoms, dtype=np.int64)
    crystal = np.repeat(np.arange(len(nat), dtype=np.int64), nat)
    # one integer key (crystal, symbol rank < 128) and a STABLE sort: equal symbols keep their original order
    return np.argsort(crystal * 128 + _SYMBOL_RANK[a], kind="stable")


def state_to_atoms(atom_types: np.ndarray, frac_coords: np.ndarray, lattices: np.ndarray,
                   natoms: Sequence[int]) -> List:
    a = np.where(atom_types <= 103, atom_types, 0)
    a = np.where(a >= 0, a, 0).astype(np.int64)
    out = []
    off = 0
    lat = np.asarray(lattices).reshape(-1, 3, 3)
    if HAVE_ASE:  # pragma: no cover
        for i, n in enumerate(natoms):
            at = _AseAtoms(numbers=a[off:off + n], cell=lat[i], pbc=True)
            at.set_scaled_positions(frac_coords[off:off + n])
            out.append(_ase_sort(at))
            off += n
        return out
    order = batch_symbol_order(a, natoms)
    z_sorted = a[order]
    x_sorted = np.asarray(frac_coords)[order].astype(np.float64)
    lat = lat.astype(np.float64)
    make = AtomsLite._from_views
    for i, n in enumerate(natoms):
        out.append(make(z_sorted[off:off + n], lat[i], x_sorted[off:off + n]))
        off += n
    return out
