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
