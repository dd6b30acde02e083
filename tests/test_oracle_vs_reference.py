"""CPU, build container only: the oracle against the LIVE unmodified reference
(imported from /root/reference through oracle/ref_shim.py)."""
import pytest
import torch

from oracle import chemeleon_oracle as O
from oracle import ref_shim
from chemeleon_b200.config import SamplerConfig
from chemeleon_b200.weights import random_init_state_dict

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref_model():
    m = ref_shim.build_reference_model(0)
    sd = random_init_state_dict(SamplerConfig(), seed=11)
    m.load_state_dict(sd, strict=False)
    return m


def test_cspnet_forward_matches_reference(ref_model):
    m = ref_model
    w = dict(m.state_dict())
    natoms = [4, 7, 5, 1]
    nat = torch.tensor(natoms)
    B, N = len(natoms), sum(natoms)
    bi = torch.arange(B).repeat_interleave(nat)
    g = torch.Generator().manual_seed(3)
    a = torch.randint(0, 104, (N,), generator=g)
    x = torch.rand(N, 3, generator=g)
    l = torch.randn(B, 3, 3, generator=g)
    temb = m.time_embed(torch.full((B,), 777))
    assert torch.equal(temb, O.time_embedding(torch.full((B,), 777), 128))
    text = torch.randn(B, 512, generator=g)
    with torch.no_grad():
        r = m.decoder(a, x, l, nat, bi, t=temb, text_embeds=text)
        o = O.cspnet_forward(w, a, x, l, nat, bi, temb, text)
    for rr, oo in zip(r, o):
        assert (rr - oo).abs().max() <= 1e-5 * rr.abs().max()


def test_d3pm_and_schedules_match_reference(ref_model):
    m = ref_model
    tab = O.D3PMTables(O.beta_tables(1000)["betas"])
    assert torch.equal(tab.q_mats, m.d3pm.q_mats)
    assert torch.equal(tab.q_one, m.d3pm.q_one_step_mats)
    assert torch.equal(O.sigma_tables(1000), m.sigma_scheduler.sigmas)
    bt = O.beta_tables(1000)
    for k in ("betas", "alphas", "alphas_cumprod", "sigmas"):
        assert torch.equal(bt[k], getattr(m.beta_scheduler, k))
    g = torch.Generator().manual_seed(5)
    N = 40
    logits = torch.randn(N, 104, generator=g) * 3
    a = torch.randint(0, 104, (N,), generator=g)
    a[::3] = 0
    u = torch.rand(N, 104, generator=g)
    for t in (1000, 500, 2, 1):
        tt = torch.full((N,), t)
        assert torch.equal(m.d3pm.p_logits(logits, a, tt, u), O.d3pm_p_logits(tab, logits, a, tt, u))


def test_product_schedules_match_reference(ref_model):
    from chemeleon_b200 import schedules as S

    m = ref_model
    bb = S.beta_buffers(1000)
    for k in ("betas", "alphas", "alphas_cumprod", "sigmas"):
        assert torch.equal(bb[k], getattr(m.beta_scheduler, k))
    assert torch.equal(S.sigma_buffer(1000), m.sigma_scheduler.sigmas)
    tt = S.time_embedding_table(1000, 128)
    assert torch.equal(tt[[0, 1, 500, 1000]], m.time_embed(torch.tensor([0, 1, 500, 1000])))
    tab = S.coefficient_table(1000, m.sigma_scheduler.sigmas_norm, 1e-5,
                              q_mats=m.d3pm.q_mats, q_one_step_mats=m.d3pm.q_one_step_mats)
    tab2 = S.coefficient_table(1000, m.sigma_scheduler.sigmas_norm, 1e-5)
    assert torch.allclose(tab, tab2, rtol=2e-6, atol=1e-9)
    t = 437
    assert float(tab[t, S.C_D_OMB]) == float(m.d3pm.q_one_step_mats[t - 1, 5, 5])
    assert float(tab[t, S.C_D_DIAG]) == float(m.d3pm.q_mats[t - 2, 5, 5])
    assert float(tab[t, S.C_D_OFF]) == float(m.d3pm.q_mats[t - 2, 5, 0])


def test_sampler_loop_matches_reference(ref_model):
    m = ref_model
    ref = ref_shim.load_reference()
    natoms = [4, 7, 5]
    B, N = 3, 16
    g = torch.Generator().manual_seed(3)
    text = torch.randn(B, 512, generator=g)
    null = torch.randn(1, 512, generator=g)
    m.text_encoder.cond, m.text_encoder.null = text, null
    rec = {}
    Orig = ref.schema.TrajectoryContainer

    class Recording(Orig):
        def __setitem__(self, t, step):
            rec[t] = (step.atom_types.clone(), step.frac_coords.clone(), step.lattices.clone())
            super().__setitem__(t, step)

    ref.chemeleon.TrajectoryContainer = Recording
    try:
        torch.manual_seed(7)
        gen = m._sample_generator(natoms, ["x"] * B, 2.0, 1e-5)
        for k, _ in enumerate(gen):
            if k == 3:
                break
    finally:
        ref.chemeleon.TrajectoryContainer = Orig
    so = O.SamplerOracle(dict(m.state_dict()), m.sigma_scheduler.sigmas_norm)
    (a, x, l), _ = so.sample(natoms, text, null, O.ReferenceNoise(7, B, N), t_stop=996)
    assert torch.equal(a, rec[996][0])
    assert (x - rec[996][1]).abs().max() < 1e-5
    assert (l - rec[996][2]).abs().max() <= 1e-5 * rec[996][2].abs().max()
