"""`chemeleon-b200 sample prompt|composition`: the reference CLI's sampling commands
(chemeleon/cli.py:20-199, scripts/sample_prompt.py:11-43, sample_target_composition.py:12-79)
on the B200 sampler.  Post-processing that needs pymatgen/SMACT (StructureMatcher dedup, CIF
export of pymatgen structures) is out of scope; structures are written as extended-XYZ-like text
unless `ase` is installed, in which case CIF is written like the reference does.
"""
from __future__ import annotations

import os
from pathlib import Path

import click


def _save(atoms_list, save_dir: Path, prefix: str = "gen") -> None:
    save_dir.mkdir(parents=True, exist_ok=True)
    for i, at in enumerate(atoms_list):
        try:
            from ase.io import write  # pragma: no cover

            write(str(save_dir / f"{prefix}_{i}.cif"), at)
        except Exception:
            with open(save_dir / f"{prefix}_{i}.txt", "w") as f:
                f.write(f"# cell (rows)\n{at.get_cell()}\n# Z, scaled position\n")
                for z, p in zip(at.get_atomic_numbers(), at.get_scaled_positions()):
                    f.write(f"{int(z)} {p[0]:.6f} {p[1]:.6f} {p[2]:.6f}\n")


def _load(kind: str, checkpoint_dir, precision):
    from .sampler import ChemeleonB200

    loader = ChemeleonB200.load_general_text_model if kind == "general" else ChemeleonB200.load_composition_model
    return loader(checkpoint_dir, precision=precision)


@click.group()
def cli():
    """B200-native Chemeleon sampler."""


@cli.group()
def sample():
    """Sample crystal structures."""


@sample.command("prompt")
@click.option("-t", "--text-input", default="A Crystal structure of LiMnO4 with orthorhombic symmetry")
@click.option("--n-samples", default=3, type=int)
@click.option("--n-atoms", default=6, type=int)
@click.option("-s", "--save-dir", default="results/prompt")
@click.option("--checkpoint-dir", default=None, help="directory holding the reference's .ckpt files")
@click.option("--precision", default="tc", type=click.Choice(["tc", "fp32"]))
def sample_prompt(text_input, n_samples, n_atoms, save_dir, checkpoint_dir, precision):
    model = _load("general", checkpoint_dir, precision)
    click.echo(f"Sampling {n_samples} structures for {text_input} with {n_atoms} atoms...")
    _save(model.sample(text_input=text_input, n_atoms=n_atoms, n_samples=n_samples), Path(save_dir))
    click.echo(f"Results saved in {save_dir}")


@sample.command("composition")
@click.option("-t", "--target-composition", default="TiO2")
@click.option("--n-samples", default=100, type=int)
@click.option("--max-natoms", default=40, type=int)
@click.option("--max-factor", default=13, type=int)
@click.option("--reduced-natoms", default=None, type=int, help="atoms in the reduced formula (default: from the formula)")
@click.option("-s", "--save-dir", default="results/composition")
@click.option("--checkpoint-dir", default=None)
@click.option("--precision", default="tc", type=click.Choice(["tc", "fp32"]))
def sample_composition(target_composition, n_samples, max_natoms, max_factor, reduced_natoms, save_dir,
                       checkpoint_dir, precision):
    """All Z-factor buckets are sampled as ONE ragged batch (the reference runs them one after the other)."""
    if reduced_natoms is None:
        from .validity import reduced_formula_counts

        reduced_natoms = sum(reduced_formula_counts(target_composition))
    model = _load("composition", checkpoint_dir, precision)
    natoms = [reduced_natoms * f for f in range(1, max_factor + 1) if reduced_natoms * f <= max_natoms
              for _ in range(n_samples)]
    click.echo(f"Sampling {len(natoms)} structures for {target_composition} in one ragged batch...")
    # validity check of the reference (lattice <= 60 A, reduced composition == target;
    # sample_target_composition.py:57-62) on the device
    valid, flags = model.sample_batch_valid(natoms, [target_composition] * len(natoms),
                                            target_composition=target_composition, min_distance=0.0)
    click.echo(f"{len(valid)} of {len(natoms)} structures pass the lattice-length / composition filter")
    _save(valid, Path(save_dir))
    click.echo(f"Results saved in {save_dir}")


if __name__ == "__main__":
    cli()
