"""CPU: the CLI surface and the bucketed composition / chemical-system driver logic
(reference: chemeleon/cli.py:7-203, sample_target_composition.py:27-41, navigate_chemical_system.py:33-60)."""
import pytest
from click.testing import CliRunner


def test_composition_prompt_is_the_alphabetical_reduced_formula():
    from chemeleon_b200.driver import alphabetical_formula, composition_prompt, reduce_counts
    from chemeleon_b200.validity import parse_formula

    assert composition_prompt("TiO2")[0] == "O2 Ti1"          # Composition("TiO2").reduced_composition.alphabetical_formula
    assert composition_prompt("Ti2O4")[0] == "O2 Ti1"
    assert composition_prompt("LiMnO4")[0] == "Li1 Mn1 O4"
    assert composition_prompt("Zn2 Ti O4")[0] == "O4 Ti1 Zn2"
    assert alphabetical_formula(reduce_counts(parse_formula("Mn2O3"))) == "Mn2 O3"


def test_plan_buckets_follows_the_reference_loops():
    from chemeleon_b200.driver import plan_buckets, reduce_counts
    from chemeleon_b200.validity import parse_formula

    red = [reduce_counts(parse_formula("TiO2")), reduce_counts(parse_formula("LiMnO4"))]
    natoms, prompt_of, buckets = plan_buckets(red, n_samples=100, max_natoms=40, max_factor=13)
    # TiO2: 3, 6, ..., 39 (13 buckets, BASELINE config 2); LiMnO4: 6, 12, ..., 36 (6 buckets)
    assert [b.n_atoms for b in buckets if b.composition == 0] == list(range(3, 40, 3))
    assert [b.n_atoms for b in buckets if b.composition == 1] == [6, 12, 18, 24, 30, 36]
    assert len(natoms) == 1900 and prompt_of[:1300] == [0] * 1300 and prompt_of[1300:] == [1] * 600
    assert buckets[13].start == 1300 and all(b.n_samples == 100 for b in buckets)
    natoms2, _, b2 = plan_buckets(red[:1], 7, max_natoms=40, max_factor=4)
    assert natoms2 == [3] * 7 + [6] * 7 + [9] * 7 + [12] * 7 and len(b2) == 4


def test_enumerate_system_dedupes_reduced_compositions():
    from chemeleon_b200.driver import alphabetical_formula, enumerate_system

    comps, screened = enumerate_system(["Ti", "O"], max_stoich=2, use_smact=False)
    names = sorted(alphabetical_formula(c) for c in comps)
    # (0,1),(0,2)->O1; (1,0),(2,0)->Ti1; (1,1),(2,2)->O1 Ti1; (1,2)->O2 Ti1; (2,1)->O1 Ti2
    assert names == ["O1", "O1 Ti1", "O1 Ti2", "O2 Ti1", "Ti1"] and not screened
    with pytest.raises(ValueError):
        enumerate_system(["Ti", "Xx"], 2)


def test_cli_surface_matches_the_reference():
    from chemeleon_b200.cli import cli

    r = CliRunner()
    for args, must in ((["sample", "prompt", "--help"], ["--text-input", "--n-samples", "--n-atoms", "--save-dir"]),
                       (["sample", "composition", "--help"], ["--target-composition", "--n-samples", "--max-natoms",
                                                              "--max-factor", "--save-dir"]),
                       (["navigate", "system", "--help"], ["--elements", "--n-samples", "--max-stoich", "--max-natoms",
                                                           "--max-factor", "--save-dir"])):
        out = r.invoke(cli, args)
        assert out.exit_code == 0, out.output
        for m in must:
            assert m in out.output


def test_cli_fails_fast_without_a_text_source(monkeypatch):
    """No --prompt-embeds and no reference package: stop with a clear message before loading weights."""
    import chemeleon_b200.cli as C

    monkeypatch.setattr(C, "_reference_text_encoder_available", lambda: False)
    out = CliRunner().invoke(C.cli, ["sample", "prompt", "-t", "LiMnO4", "--checkpoint-dir", "/nonexistent"])
    assert out.exit_code != 0 and "--prompt-embeds" in out.output and "Traceback" not in out.output


def test_cli_composition_uses_the_trained_prompt_format(monkeypatch, tmp_path):
    """`sample composition -t TiO2` conditions on 'O2 Ti1' and issues all 13 buckets as one batch."""
    import chemeleon_b200.cli as C
    import chemeleon_b200.driver as D

    seen = {}

    class FakeModel:
        text_guide = True

    def fake_sample_compositions(model, formulas, n_samples, max_natoms, max_factor, **kw):
        red = [D.reduce_counts(D.parse_formula(f)) for f in formulas]
        natoms, prompt_of, buckets = D.plan_buckets(red, n_samples, max_natoms, max_factor)
        seen.update(prompts=[D.alphabetical_formula(c) for c in red], natoms=natoms, buckets=buckets)
        return dict(valid={0: []}, natoms=natoms, buckets=buckets)

    monkeypatch.setattr(C, "_load", lambda *a, **k: FakeModel())
    monkeypatch.setattr(D, "sample_compositions", fake_sample_compositions)
    out = CliRunner().invoke(C.cli, ["sample", "composition", "-t", "TiO2", "--n-samples", "100", "-s", str(tmp_path)])
    assert out.exit_code == 0, out.output
    assert "target composition: O2 Ti1" in out.output
    assert seen["prompts"] == ["O2 Ti1"] and len(seen["buckets"]) == 13 and len(seen["natoms"]) == 1300
