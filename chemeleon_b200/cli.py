"""`chemeleon-b200`: the reference CLI's commands on the B200 sampler
(chemeleon/cli.py:7-203; scripts/sample_prompt.py:11-43, sample_target_composition.py:12-79,
navigate_chemical_system.py:15-103) with the same options and defaults:

    sample prompt       -t/--text-input --n-samples --n-atoms -s/--save-dir
    sample composition  -t/--target-composition --n-samples --max-natoms --max-factor -s/--save-dir
    navigate system     -e/--elements --n-samples --max-stoich --max-natoms --max-factor -s/--save-dir

Text conditioning: the checkpoint's own text tail (text_encoder.text_emb.*, null embedding) runs on
the GPU; what has to come from outside is the language-model embedding of each distinct prompt --
either `--prompt-embeds FILE` (torch file, dict prompt -> [embed_dim] tensor), or the reference
package importable next to this one (its `TextEncoder.text_encode` is then called once per distinct
prompt).  Without either the command stops BEFORE loading any weights.

Post-processing that needs pymatgen / SMACT (StructureMatcher de-duplication, charge-neutrality
screening) is out of scope; structures are written as CIF when `ase` is importable, else as plain text.
Run under `torchrun` to shard the batch over the GPUs of the box.
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


def _reference_text_encoder_available() -> bool:
    try:
        import chemeleon.text_encoder.text_encoder  # noqa: F401

        return True
    except Exception:
        return False


def _check_text_source(prompt_embeds) -> None:
    """Fail fast, before any weight is loaded."""
    if prompt_embeds is None and not _reference_text_encoder_available():
        raise click.UsageError(
            "text prompts need the language-model embedding of each prompt: pass --prompt-embeds FILE (torch file, "
            "dict prompt -> [embed_dim] tensor) or install the reference package next to this one (its TextEncoder / "
            "BERT + CrystalClip weights are then used for `text_encode`)")


def _load(kind: str, checkpoint_dir, precision, prompt_embeds=None):
    import torch

    from .sampler import ChemeleonB200

    _check_text_source(prompt_embeds)
    loader = ChemeleonB200.load_general_text_model if kind == "general" else ChemeleonB200.load_composition_model
    rank = int(os.environ.get("LOCAL_RANK", "0"))
    model = loader(checkpoint_dir, precision=precision, device=f"cuda:{rank}")
    if not model.engine.has_text_tail:
        raise click.UsageError("the checkpoint holds no text_encoder.text_emb.* weights (not a text-guided model?)")
    if prompt_embeds is not None:
        for prompt, emb in torch.load(prompt_embeds, map_location="cpu").items():
            model.set_prompt_embedding(prompt, emb)
    else:  # pragma: no cover - needs the reference package and its language-model weights
        from chemeleon.text_encoder.text_encoder import TextEncoder

        hp = model.source_hparams
        model.text_encoder = TextEncoder(text_encoder_name=hp.get("text_encoder", "lfoppiano/MatTPUSciBERT"),
                                         text_embed_dim=hp.get("text_embed_dim", 768),
                                         max_text_len=hp.get("max_text_len", 256), text_dim=hp.get("text_dim", 512))
    return model


def _maybe_init_distributed():
    import torch.distributed as dist

    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        import torch

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    return int(os.environ.get("RANK", "0"))


COMMON = [
    click.option("--checkpoint-dir", default=None, help="directory holding the reference's .ckpt files"),
    click.option("--precision", default="tc", type=click.Choice(["tc", "fp32"])),
    click.option("--prompt-embeds", default=None, type=click.Path(exists=True, dir_okay=False),
                 help="torch file: dict prompt -> language-model embedding [embed_dim]"),
]


def _common(f):
    for opt in reversed(COMMON):
        f = opt(f)
    return f


@click.group(help="B200-native Chemeleon sampler (reference CLI surface).")
def cli():
    pass


@cli.group(help="Sample crystal structures.")
def sample():
    pass


@cli.group(help="Commands related to chemical system navigation.")
def navigate():
    pass


@sample.command("prompt")
@click.option("-t", "--text-input", default="A Crystal structure of LiMnO4 with orthorhombic symmetry", show_default=True)
@click.option("--n-samples", default=3, type=int, show_default=True)
@click.option("--n-atoms", default=6, type=int, show_default=True)
@click.option("-s", "--save-dir", default="results/prompt", show_default=True)
@_common
def sample_prompt(text_input, n_samples, n_atoms, save_dir, checkpoint_dir, precision, prompt_embeds):
    """scripts/sample_prompt.py:11-43."""
    model = _load("general", checkpoint_dir, precision, prompt_embeds)
    click.echo(f"Sampling {n_samples} structures for {text_input} with {n_atoms} atoms...")
    _save(model.sample(text_input=text_input, n_atoms=n_atoms, n_samples=n_samples), Path(save_dir))
    click.echo(f"Results saved in {save_dir}")


@sample.command("composition")
@click.option("-t", "--target-composition", default="TiO2", show_default=True)
@click.option("--n-samples", default=100, type=int, show_default=True)
@click.option("--max-natoms", default=40, type=int, show_default=True)
@click.option("--max-factor", default=13, type=int, show_default=True)
@click.option("-s", "--save-dir", default="results/composition", show_default=True)
@_common
def sample_composition(target_composition, n_samples, max_natoms, max_factor, save_dir, checkpoint_dir, precision,
                       prompt_embeds):
    """scripts/sample_target_composition.py:12-79 -- all Z-factor buckets as ONE ragged batch."""
    from .driver import composition_prompt, sample_compositions

    prompt, _ = composition_prompt(target_composition)      # 'TiO2' -> 'O2 Ti1', the trained prompt format
    click.echo(f"target composition: {prompt}")
    rank = _maybe_init_distributed()
    model = _load("composition", checkpoint_dir, precision, prompt_embeds)
    res = sample_compositions(model, [target_composition], n_samples, max_natoms, max_factor)
    valid = res["valid"][0]
    if rank == 0:
        click.echo(f"Sampled {len(res['natoms'])} structures in {len(res['buckets'])} buckets as one ragged batch; "
                   f"{len(valid)} pass the lattice-length / composition filter")
        _save(valid, Path(save_dir), prefix=f"gen_{prompt.replace(' ', '')}")
        click.echo(f"Results saved in {save_dir}")


@navigate.command("system")
@click.option("-e", "--elements", default="Zn,Ti,O", show_default=True,
              help="Comma-separated list of elements to navigate the chemical system. e.g. Zn,Ti,O")
@click.option("--n-samples", default=100, type=int, show_default=True)
@click.option("--max-stoich", default=8, type=int, show_default=True)
@click.option("--max-natoms", default=40, type=int, show_default=True)
@click.option("--max-factor", default=13, type=int, show_default=True)
@click.option("-s", "--save-dir", default="results/navigate", show_default=True)
@click.option("--chunk", default=64, type=int, show_default=True, help="compositions per ragged batch")
@_common
def navigate_system(elements, n_samples, max_stoich, max_natoms, max_factor, save_dir, chunk, checkpoint_dir,
                    precision, prompt_embeds):
    """scripts/navigate_chemical_system.py:15-103 -- compositions x Z factors as ragged multi-GPU batches."""
    from .driver import alphabetical_formula, enumerate_system, sample_compositions

    els = [e.strip() for e in elements.split(",") if e.strip()]
    comps, screened = enumerate_system(els, max_stoich)
    click.echo(f"Number of unique {'valid ' if screened else ''}compositions: {len(comps)}"
               + ("" if screened else " (SMACT not importable: no charge-neutrality screen)"))
    rank = _maybe_init_distributed()
    model = _load("composition", checkpoint_dir, precision, prompt_embeds)
    total = 0
    for c0 in range(0, len(comps), chunk):
        res = sample_compositions(model, comps[c0:c0 + chunk], n_samples, max_natoms, max_factor, seed=c0)
        if rank == 0:
            for ci, atoms in res["valid"].items():
                total += len(atoms)
                _save(atoms, Path(save_dir), prefix=f"gen_{alphabetical_formula(comps[c0 + ci]).replace(' ', '')}")
    if rank == 0:
        click.echo(f"{total} structures pass the filters; results saved in {save_dir}")


if __name__ == "__main__":
    cli()
