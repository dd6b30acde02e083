"""Size-independent properties of the CUDA decoder and the sampler API:
permutation equivariance over the atoms of a crystal, invariance to integer shifts of the
fractional coordinates, independence of a crystal's outputs from its batch-mates, and the
`sample(stream=..., return_trajectory=...)` contract (chemeleon.py:469-490)."""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
PRECISIONS = ["fp32", "tc"]
TOL = {"fp32": 2e-5, "tc": 1e-3}


@pytest.fixture(scope="module")
def nets():
    from chemeleon_b200.config import SamplerConfig
    from chemeleon_b200.cspnet import CSPNetB200
    from chemeleon_b200.weights import random_init_state_dict

    cfg = SamplerConfig(num_layers=2)
    sd = random_init_state_dict(cfg, seed=7, perturb_ln=True)
    return {p: CSPNetB200(sd, cfg, precision=p) for p in PRECISIONS}


def _inputs(natoms, seed=0):
    g = torch.Generator().manual_seed(seed)
    nat = torch.tensor(natoms)
    B, N = len(natoms), sum(natoms)
    bi = torch.arange(B).repeat_interleave(nat)
    a = torch.randint(1, 104, (N,), generator=g)
    x = torch.rand(N, 3, generator=g)
    l = torch.randn(B, 3, 3, generator=g) + 3 * torch.eye(3)
    temb = torch.randn(1, 128, generator=g).repeat(B, 1)
    text = torch.randn(B, 512, generator=g)
    return a, x, l, nat, bi, temb, text


@pytest.mark.parametrize("precision", PRECISIONS)
def test_permutation_equivariance(nets, precision):
    net = nets[precision]
    natoms = [7, 20, 5]
    a, x, l, nat, bi, temb, text = _inputs(natoms, 1)
    out = net(a, x, l, nat, bi, t=temb, text_embeds=text)
    perm = torch.arange(sum(natoms))
    g = torch.Generator().manual_seed(5)
    off = 0
    for n in natoms:                                   # shuffle the atoms inside every crystal
        perm[off:off + n] = off + torch.randperm(n, generator=g)
        off += n
    outp = net(a[perm], x[perm], l, nat, bi, t=temb, text_embeds=text)
    tol = TOL[precision]
    assert rel_err(outp.atom_types_out.cpu(), out.atom_types_out.cpu()[perm]) < tol
    assert rel_err(outp.coords_out.cpu(), out.coords_out.cpu()[perm]) < tol
    assert rel_err(outp.lattice_out.cpu(), out.lattice_out.cpu()) < tol


@pytest.mark.parametrize("precision", PRECISIONS)
def test_integer_shift_invariance(nets, precision):
    """Only (x_j - x_i) mod 1 enters the decoder (cspnet.py:324): shifting atoms by lattice vectors changes nothing."""
    net = nets[precision]
    natoms = [6, 13]
    a, x, l, nat, bi, temb, text = _inputs(natoms, 2)
    out = net(a, x, l, nat, bi, t=temb, text_embeds=text)
    g = torch.Generator().manual_seed(9)
    shift = torch.randint(-3, 4, x.shape, generator=g).float()
    outs = net(a, x + shift, l, nat, bi, t=temb, text_embeds=text)
    tol = 5 * TOL[precision]                          # x + k is rounded to fp32: the wrapped difference moves by ~1e-6
    for o, r in zip(outs[:3], out[:3]):
        assert rel_err(o.cpu(), r.cpu()) < tol


@pytest.mark.parametrize("precision", PRECISIONS)
def test_batch_composition_invariance(nets, precision):
    """A crystal's outputs do not depend on its batch-mates (the property sharding by sample relies on)."""
    net = nets[precision]
    natoms = [9, 20, 4, 20, 33]
    a, x, l, nat, bi, temb, text = _inputs(natoms, 3)
    out = net(a, x, l, nat, bi, t=temb, text_embeds=text)
    off = [0]
    for n in natoms:
        off.append(off[-1] + n)
    for b in (1, 2, 4):
        sl = slice(off[b], off[b + 1])
        solo = net(a[sl], x[sl], l[b:b + 1], nat[b:b + 1], torch.zeros(natoms[b], dtype=torch.long),
                   t=temb[b:b + 1], text_embeds=text[b:b + 1])
        tol = TOL[precision]
        assert rel_err(solo.atom_types_out.cpu(), out.atom_types_out.cpu()[sl]) < tol
        assert rel_err(solo.coords_out.cpu(), out.coords_out.cpu()[sl]) < tol
        assert rel_err(solo.lattice_out.cpu(), out.lattice_out.cpu()[b:b + 1]) < tol


def test_sample_api_contract():
    """sample(): list of structures; stream=True: generator with one list per timestep;
    return_trajectory=True: the same lists materialised; the last step equals the plain result."""
    from chemeleon_b200.config import SamplerConfig
    from chemeleon_b200.sampler import ChemeleonB200
    from chemeleon_b200.weights import random_init_state_dict

    cfg = SamplerConfig(timesteps=6)
    model = ChemeleonB200(random_init_state_dict(cfg, seed=0, head_scale=0.01, lattice_identity=True), cfg)
    g = torch.Generator().manual_seed(0)
    emb = dict(text_embeds=torch.randn(3, cfg.text_dim, generator=g), null_text_embeds=torch.randn(1, cfg.text_dim, generator=g))
    final = model.sample_batch([5, 5, 5], **emb, seed=11)
    assert len(final) == 3 and all(len(at) == 5 for at in final)
    gen = model._sample_generator([5, 5, 5], None, 2.0, 1e-5, **emb, seed=11)
    assert hasattr(gen, "__next__")
    steps = list(gen)
    assert len(steps) == cfg.timesteps and all(len(s) == 3 for s in steps)
    for at_f, at_s in zip(final, steps[-1]):
        assert (at_f.get_atomic_numbers() == at_s.get_atomic_numbers()).all()
        assert abs(at_f.get_scaled_positions() - at_s.get_scaled_positions()).max() < 1e-6


def test_text_tail_on_device_matches_reference():
    """cb2_text_condition (text_emb MLP + learned null embedding + the text half of mlp_cond, once per
    DISTINCT prompt) against the reference's real TextEncoder.get_text_embeds (fixture), and
    sample(text_input=...) running from a registered language-model embedding alone."""
    from helpers import load_golden
    from chemeleon_b200.config import SamplerConfig
    from chemeleon_b200.sampler import ChemeleonB200
    from chemeleon_b200.weights import random_init_state_dict

    g = load_golden("text_tail")
    cfg = SamplerConfig(timesteps=4)
    sd = random_init_state_dict(cfg, seed=int(g["weight_seed"]), head_scale=0.01, lattice_identity=True, text_tail_dim=768)
    model = ChemeleonB200(sd, cfg)
    assert model.engine.has_text_tail
    enc, ids = torch.from_numpy(g["enc"]), g["prompt_ids"].tolist()
    cond = model.condition_from_encoder(enc, ids)
    P = enc.shape[0]
    assert cond.rows.shape == (P + 1, 1024) and cond.row_of.tolist() == ids + [P] * len(ids)
    wc = sd["decoder.film_layer.mlp_cond.0.weight"].double()
    bc = sd["decoder.film_layer.mlp_cond.0.bias"].double()
    want_cond = torch.from_numpy(g["cond"]).double() @ wc[:, 128:].t() + bc      # reference embeddings, folded in fp64
    want_null = torch.from_numpy(g["null"]).double() @ wc[:, 128:].t() + bc
    rows = cond.rows.cpu().double()
    assert rel_err(rows[torch.tensor(ids)], want_cond) < 1e-5
    assert rel_err(rows[P:P + 1].expand(len(ids), -1), want_null) < 1e-5
    # the headline API from a precomputed language-model embedding: no text encoder object at all
    model.set_prompt_embedding("LiMnO4 orthorhombic", enc[2])
    atoms = model.sample("LiMnO4 orthorhombic", 6, 3, seed=5)
    assert len(atoms) == 3 and all(len(a) == 6 for a in atoms)
    direct = model.sample_batch([6] * 3, text_embeds=model.condition_from_encoder(enc[2:3], [0, 0, 0]), seed=5)
    for a, b in zip(atoms, direct):
        assert (a.get_atomic_numbers() == b.get_atomic_numbers()).all()
        assert abs(a.get_scaled_positions() - b.get_scaled_positions()).max() == 0.0
    with pytest.raises(ValueError):
        model.sample("an unknown prompt", 6, 1)


def test_driver_samples_all_buckets_as_one_batch():
    """driver.sample_compositions: every (composition, Z factor) bucket in ONE ragged batch through
    dist.sample_sharded, text tail once per distinct prompt, validity flags from the device filter;
    per-crystal results equal the plain single-call sampler (sharding / batch order invariance)."""
    from chemeleon_b200 import driver
    from chemeleon_b200.config import SamplerConfig
    from chemeleon_b200.sampler import ChemeleonB200
    from chemeleon_b200.weights import random_init_state_dict

    cfg = SamplerConfig(timesteps=5)
    sd = random_init_state_dict(cfg, seed=3, head_scale=0.01, lattice_identity=True, text_tail_dim=768)
    model = ChemeleonB200(sd, cfg)
    g = torch.Generator().manual_seed(0)
    for p in ("O2 Ti1", "Li1 Mn1 O4"):
        model.set_prompt_embedding(p, torch.randn(768, generator=g))
    res = driver.sample_compositions(model, ["TiO2", "LiMnO4"], n_samples=3, max_natoms=13, max_factor=13, seed=9)
    assert res["prompts"] == ["O2 Ti1", "Li1 Mn1 O4"]
    assert [b.n_atoms for b in res["buckets"]] == [3, 6, 9, 12, 6, 12]
    natoms = res["natoms"]
    assert natoms == [3] * 3 + [6] * 3 + [9] * 3 + [12] * 3 + [6] * 3 + [12] * 3
    a, x, l = res["state"]
    assert a.numel() == sum(natoms) and l.shape[0] == len(natoms) and res["flags"].numel() == len(natoms)
    assert len(res["atoms"]) == len(natoms) and all(len(at) == n for at, n in zip(res["atoms"], natoms))
    assert sum(len(v) for v in res["valid"].values()) == int((res["flags"] == 0).sum())
    # same batch through the plain sampler entry point
    cond, _ = model._embed_texts(["O2 Ti1"] * 12 + ["Li1 Mn1 O4"] * 6)
    a2, x2, l2 = model.sample_states(natoms, cond, None, seed=9)
    assert torch.equal(a2, a)
    assert torch.allclose(x2, x, atol=1e-5) and torch.allclose(l2.reshape(-1, 9), l.reshape(-1, 9), atol=1e-5, rtol=1e-5)


def test_streaming_frames_match_the_atoms_generator():
    """stream=True: frames packed on the device and copied on a side stream (wire format of
    include/chemeleon_b200.h) carry exactly what the List[Atoms] generator yields, one per timestep."""
    from chemeleon_b200.config import SamplerConfig
    from chemeleon_b200.sampler import ChemeleonB200
    from chemeleon_b200.streaming import Frame
    from chemeleon_b200.weights import random_init_state_dict

    cfg = SamplerConfig(timesteps=9)
    model = ChemeleonB200(random_init_state_dict(cfg, seed=0, head_scale=0.01, lattice_identity=True), cfg)
    g = torch.Generator().manual_seed(0)
    natoms = [5, 3, 7]
    emb = dict(text_embeds=torch.randn(3, cfg.text_dim, generator=g), null_text_embeds=torch.randn(1, cfg.text_dim, generator=g))
    frames = list(model._sample_generator(natoms, None, 2.0, 1e-5, **emb, seed=4, frames=True, depth=3))
    atoms = list(model._sample_generator(natoms, None, 2.0, 1e-5, **emb, seed=4, depth=2))
    assert len(frames) == cfg.timesteps == len(atoms)
    assert [f.t for f in frames] == list(range(cfg.timesteps - 1, -1, -1))      # index of the state, like trajectory[t-1]
    for f, ats in zip(frames, atoms):
        assert f.n_nodes == 15 and f.n_graphs == 3 and len(f.to_bytes()) == 16 + 16 + 15 * 12 + 3 * 36
        back = Frame.from_bytes(f.to_bytes()).to_atoms(natoms)
        for a, b in zip(back, ats):
            assert (a.get_atomic_numbers() == b.get_atomic_numbers()).all()
            assert abs(a.get_scaled_positions() - b.get_scaled_positions()).max() == 0.0
            assert abs(a.get_cell() - b.get_cell()).max() == 0.0
    # the last frame is the final state of the plain call
    final = model.sample_batch(natoms, **emb, seed=4)
    for a, b in zip(frames[-1].to_atoms(natoms), final):
        assert (a.get_atomic_numbers() == b.get_atomic_numbers()).all()
        assert abs(a.get_scaled_positions() - b.get_scaled_positions()).max() < 1e-6
