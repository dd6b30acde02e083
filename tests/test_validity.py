"""Validity pre-filter: oracle known answers on the CPU, CUDA kernel
against the oracle on the GPU (flags bit-exact, distances to 1e-5 relative)."""
import math

import numpy as np
import pytest
import torch

from oracle import validity_oracle as VO


def test_formula_parsing_and_reduction():
    from chemeleon_b200.validity import parse_formula, reduced_formula_counts

    c = parse_formula("TiO2")
    assert c[22] == 1 and c[8] == 2 and sum(c) == 3
    c = reduced_formula_counts("Ti4 O8")
    assert c[22] == 1 and c[8] == 2
    c = reduced_formula_counts("LiMnO4")
    assert c[3] == 1 and c[25] == 1 and c[8] == 4
    with pytest.raises(ValueError):
        parse_formula("Ti(OH)2")
    with pytest.raises(ValueError):
        parse_formula("Xx2")


def test_oracle_known_answers():
    cubic = np.eye(3) * 4.0
    # two atoms 0.4 A apart through the periodic boundary
    r = VO.validity_flags([8, 8], [[0.02, 0.5, 0.5], [0.92, 0.5, 0.5]], cubic)
    assert abs(r["min_dist"] - 0.4) < 1e-12 and r["flags"] == VO.INVALID_DISTANCE
    # rock-salt-like pair at half the cell: 2 A apart, fine
    r = VO.validity_flags([11, 17], [[0, 0, 0], [0.5, 0, 0]], cubic)
    assert abs(r["min_dist"] - 2.0) < 1e-12 and r["flags"] == 0 and abs(r["max_abc"] - 4.0) < 1e-12
    # coincident atoms are skipped (dist_mat > 0), a single atom has no positive distance
    r = VO.validity_flags([1, 1, 8], [[0.1, 0.1, 0.1], [0.1, 0.1, 0.1], [0.6, 0.1, 0.1]], cubic)
    assert abs(r["min_dist"] - 2.0) < 1e-12
    assert math.isinf(VO.validity_flags([1], [[0, 0, 0]], cubic)["min_dist"])
    # long lattice vector
    assert VO.validity_flags([1, 1], [[0, 0, 0], [0.5, 0.5, 0.5]], np.diag([61.0, 5, 5]))["flags"] == VO.INVALID_LATTICE
    # strongly sheared cell: the closest image is not among the 27 neighbours of the unwrapped difference
    cell = np.array([[3.0, 0, 0], [8.7, 3.0, 0], [0, 0, 3.0]])
    r = VO.validity_flags([1, 1], [[0, 0, 0], [0.1, 0.5, 0]], cell)
    v = np.array([0.1, 0.5, 0]) @ cell
    best = min(np.linalg.norm(v + np.array([i, j, 0.0]) @ cell) for i in range(-6, 7) for j in range(-3, 4))
    assert abs(r["min_dist"] - best) < 1e-12
    # composition
    tgt = np.zeros(104, dtype=np.int64); tgt[22] = 1; tgt[8] = 2
    assert VO.validity_flags([22, 8, 8, 22, 8, 8], np.random.rand(6, 3), cubic * 3, tgt)["flags"] & VO.INVALID_COMPOSITION == 0
    assert VO.validity_flags([22, 8, 8, 22, 8, 1], np.random.rand(6, 3), cubic * 3, tgt)["flags"] & VO.INVALID_COMPOSITION


def _random_batch(seed, B=40):
    g = np.random.default_rng(seed)
    natoms = [int(v) for v in g.integers(1, 41, size=B)]
    cells, fracs, nums = [], [], []
    for b, n in enumerate(natoms):
        kind = b % 5
        if kind == 0:
            cell = np.diag(g.uniform(2.0, 12.0, size=3))
        elif kind == 1:
            cell = np.diag(g.uniform(3.0, 9.0, size=3)) + g.uniform(-2.5, 2.5, size=(3, 3))
        elif kind == 2:
            cell = np.array([[3.0, 0, 0], [g.uniform(5, 9), 3.0, 0], [0, 0, g.uniform(3, 70)]])
        elif kind == 3:
            cell = np.diag(g.uniform(0.7, 1.6, size=3)) + g.uniform(-0.2, 0.2, size=(3, 3))   # tiny cell
        else:
            cell = g.uniform(-6, 6, size=(3, 3))
        frac = g.uniform(-0.3, 1.3, size=(n, 3))
        if n > 2 and b % 7 == 0:
            frac[1] = frac[0]                                   # coincident pair
        z = g.integers(0, 104, size=n)
        if b % 3 == 0:
            z = np.array(([22, 8, 8] * n)[:n])
        cells.append(cell); fracs.append(frac); nums.append(z)
    return natoms, cells, fracs, nums


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_validity_kernel_matches_oracle(seed):
    from chemeleon_b200.validity import validity_flags

    natoms, cells, fracs, nums = _random_batch(seed)
    a = torch.from_numpy(np.concatenate(nums)).cuda()
    x = torch.from_numpy(np.concatenate(fracs)).float().cuda()
    lat = torch.from_numpy(np.stack(cells)).float().cuda()
    flags, dmin, abc = validity_flags(a, x, lat, natoms, target="TiO2")
    flags, dmin, abc = flags.cpu().numpy(), dmin.cpu().numpy(), abc.cpu().numpy()
    tgt = np.zeros(104, dtype=np.int64); tgt[22] = 1; tgt[8] = 2
    checked = 0
    for b in range(len(natoms)):
        # the oracle sees the same float32 inputs as the kernel
        ref = VO.validity_flags(nums[b], fracs[b].astype(np.float32), cells[b].astype(np.float32), tgt)
        assert abs(abc[b] - ref["max_abc"]) <= 1e-5 * ref["max_abc"]
        assert (flags[b] & VO.INVALID_COMPOSITION) == (ref["flags"] & VO.INVALID_COMPOSITION), b
        if abs(ref["max_abc"] - 60.0) > 1e-3:
            assert (flags[b] & VO.INVALID_LATTICE) == (ref["flags"] & VO.INVALID_LATTICE), b
        if math.isinf(ref["min_dist"]):
            assert math.isinf(dmin[b])
            continue
        if abs(ref["min_dist"] - 0.5) > 1e-3:                   # away from the threshold the decision is exact
            assert (flags[b] & VO.INVALID_DISTANCE) == (ref["flags"] & VO.INVALID_DISTANCE), (b, dmin[b], ref)
        if ref["min_dist"] < 0.5:                               # below the threshold the value itself is exact
            assert abs(dmin[b] - ref["min_dist"]) <= 2e-5 * max(1.0, ref["min_dist"]) + 1e-6, (b, dmin[b], ref)
            checked += 1
        else:
            assert dmin[b] >= ref["min_dist"] * (1 - 2e-5)      # never below the true minimum
    assert checked > 3


@pytest.mark.gpu
def test_validity_on_sampler_state():
    """End of a short sampling run: the filter runs on the device state without a host round trip."""
    from chemeleon_b200.config import SamplerConfig
    from chemeleon_b200.sampler import ChemeleonB200
    from chemeleon_b200.validity import validity_flags
    from chemeleon_b200.weights import random_init_state_dict

    cfg = SamplerConfig(timesteps=8)
    model = ChemeleonB200(random_init_state_dict(cfg, seed=0, head_scale=0.01, lattice_identity=True), cfg)
    natoms = [3, 6, 6, 9]
    a, x, l = model.sample_states(natoms, torch.randn(4, cfg.text_dim), torch.randn(1, cfg.text_dim), seed=3)
    flags, dmin, abc = validity_flags(a.cuda(), x.cuda(), l.cuda(), natoms)
    for b, (n0, n) in enumerate(zip(np.cumsum([0] + natoms[:-1]), natoms)):
        ref = VO.validity_flags(a[n0:n0 + n].cpu().numpy(), x[n0:n0 + n].cpu().numpy(), l[b].cpu().numpy())
        assert abs(float(abc[b]) - ref["max_abc"]) <= 1e-5 * ref["max_abc"]
        if abs(ref["min_dist"] - 0.5) > 1e-3 and abs(ref["max_abc"] - 60) > 1e-3:
            assert int(flags[b]) == ref["flags"]
