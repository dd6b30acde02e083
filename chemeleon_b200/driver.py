"""Bucketed composition / chemical-system drivers on top of the sampler.

The reference samples one (composition, Z factor) bucket after the other -- 13 sequential
1000-step `model.sample()` calls per composition (chemeleon/scripts/sample_target_composition.py:37-51)
and that again for every composition of a chemical system (navigate_chemical_system.py:49-87).
Here every bucket of every composition is ONE ragged batch: one captured 1000-step run, sharded over
the GPUs of the box when `torch.distributed` is initialised (`dist.sample_sharded`), the text tail
evaluated once per DISTINCT prompt, and the reference's validity filters (lattice <= 60 A, reduced
composition == target; sample_target_composition.py:57-62) evaluated on the device.

Out of scope (CPU chemistry tooling, not installable offline): SMACT charge-neutrality screening
(used when importable), pymatgen `StructureMatcher` de-duplication, CIF export without `ase`.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .atoms import SYMBOLS
from .validity import INVALID_COMPOSITION, INVALID_LATTICE, parse_formula, validity_flags


def reduce_counts(counts: Sequence[int]) -> Tuple[int, ...]:
    g = 0
    for c in counts:
        g = math.gcd(g, int(c))
    if g == 0:
        raise ValueError("empty composition")
    return tuple(int(c) // g for c in counts)


def alphabetical_formula(counts: Sequence[int]) -> str:
    """pymatgen's `Composition.alphabetical_formula` of integer counts per atomic number: elements
    sorted by symbol, 'El<amount>' joined by blanks ('O2 Ti1') -- the prompt format the composition
    model is trained on (datasets/dataset_utils.py:62, sample_target_composition.py:31,41)."""
    items = sorted((SYMBOLS[z], int(c)) for z, c in enumerate(counts) if c)
    return " ".join(f"{s}{c}" for s, c in items)


def composition_prompt(formula: str) -> Tuple[str, Tuple[int, ...]]:
    """'TiO2' -> ('O2 Ti1', reduced counts per atomic number)."""
    red = reduce_counts(parse_formula(formula))
    return alphabetical_formula(red), red


def enumerate_system(elements: Sequence[str], max_stoich: int = 8, use_smact: Optional[bool] = None):
    """All reduced compositions of a chemical system with stoichiometries 0..max_stoich
    (navigate_chemical_system.py:33-41), de-duplicated; screened by `smact_validity` when SMACT is
    importable (`use_smact=None`: use it if present; True: require it).  Returns
    (list of reduced counts per atomic number, whether the SMACT screen was applied)."""
    zs = []
    for el in elements:
        if el not in SYMBOLS[1:]:
            raise ValueError(f"unknown element {el!r}")
        zs.append(SYMBOLS.index(el))
    if len(set(zs)) != len(zs):
        raise ValueError("duplicate element")
    seen, out = set(), []
    for amounts in itertools.product(range(max_stoich + 1), repeat=len(zs)):
        if max(amounts) == 0:
            continue
        counts = [0] * 104
        for z, a in zip(zs, amounts):
            counts[z] = a
        red = reduce_counts(counts)
        if red not in seen:
            seen.add(red)
            out.append(red)
    screened = False
    if use_smact is not False:
        try:  # pragma: no cover - SMACT / pymatgen are absent in the offline image
            from pymatgen.core import Composition
            from smact.screening import smact_validity

            out = [c for c in out if smact_validity(Composition(alphabetical_formula(c).replace(" ", "")))]
            screened = True
        except Exception:
            if use_smact:
                raise
    return out, screened


@dataclass
class Bucket:
    composition: int      # index into the composition list
    factor: int           # Z factor
    n_atoms: int
    start: int            # first sample of the bucket in the batch
    n_samples: int


def plan_buckets(reduced: Sequence[Sequence[int]], n_samples: int, max_natoms: int, max_factor: int):
    """The (composition, Z factor) buckets of the reference's loops as one sample list, composition by
    composition, factor by factor (sample_target_composition.py:37-41)."""
    natoms: List[int] = []
    prompt_of: List[int] = []
    buckets: List[Bucket] = []
    for ci, counts in enumerate(reduced):
        base = int(sum(counts))
        for f in range(1, max_factor + 1):
            if base * f > max_natoms:
                break
            buckets.append(Bucket(ci, f, base * f, len(natoms), n_samples))
            natoms += [base * f] * n_samples
            prompt_of += [ci] * n_samples
    return natoms, prompt_of, buckets


def sample_compositions(model, formulas_or_counts: Sequence, n_samples: int = 100, max_natoms: int = 40,
                        max_factor: int = 13, cond_scale: float = 2.0, step_lr: float = 1e-5, seed: int = 0,
                        max_length: float = 60.0, t_stop: int = 0):
    """Sample every Z-factor bucket of every composition as ONE ragged batch and apply the
    reference's filters on the device.  `formulas_or_counts`: formulas ('TiO2') or reduced counts.
    Returns a dict: prompts, buckets, natoms, per-sample flags (host int32, 0 = valid), the device
    state (a, x, l) of the whole batch, and `valid`: per composition the list of valid Atoms."""
    from . import dist as cdist

    reduced = [reduce_counts(parse_formula(f)) if isinstance(f, str) else reduce_counts(f) for f in formulas_or_counts]
    prompts = [alphabetical_formula(c) for c in reduced]
    natoms, prompt_of, buckets = plan_buckets(reduced, n_samples, max_natoms, max_factor)
    if not natoms:
        raise ValueError("no bucket fits max_natoms")
    cond = None
    if model.text_guide:
        cond, _ = model._embed_texts([prompts[i] for i in prompt_of])   # text tail once per DISTINCT prompt
    a, x, l = cdist.sample_sharded(model, natoms, cond, None, cond_scale, step_lr, seed=seed, t_stop=t_stop)
    model.check_flags()
    # validity on the device, composition by composition (each has its own target)
    flags = torch.zeros(len(natoms), dtype=torch.int32, device=x.device)
    node_off = [0]
    for n in natoms:
        node_off.append(node_off[-1] + n)
    with torch.cuda.device(model.device):
        for ci, counts in enumerate(reduced):
            bs = [b for b in buckets if b.composition == ci]
            if not bs:
                continue
            s0, s1 = bs[0].start, bs[-1].start + bs[-1].n_samples
            n0, n1 = node_off[s0], node_off[s1]
            f, _, _ = validity_flags(a[n0:n1], x[n0:n1], l[s0:s1], natoms[s0:s1], alphabetical_formula(counts).replace(" ", ""),
                                     max_length, 0.0)
            flags[s0:s1] = f & (INVALID_LATTICE | INVALID_COMPOSITION)
    flags_h = flags.cpu()
    atoms = model._to_atoms(a, x, l, natoms)
    valid: Dict[int, list] = {ci: [] for ci in range(len(reduced))}
    for i, (at, fl) in enumerate(zip(atoms, flags_h.tolist())):
        if fl == 0:
            valid[prompt_of[i]].append(at)
    return dict(prompts=prompts, buckets=buckets, natoms=natoms, flags=flags_h, state=(a, x, l), valid=valid,
                atoms=atoms)
