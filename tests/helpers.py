"""Shared test helpers: golden fixtures, weights regenerated from seeds."""
import os

import numpy as np
import torch

from chemeleon_b200.config import SamplerConfig
from chemeleon_b200.weights import random_init_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["c1_tamed_1000", "c1_bounded_1000", "c1_full_6", "ragged_full_4",
                # Angstrom regime (cells of 8..25 A with shear for the whole run; |L L^T| ~ 1e2..1e3)
                "ang_c1_1000", "ang_n20_6", "ang_n40_4"]
FORWARD_GOLDEN = "ang_forward"


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def weight_checksum(sd):
    keys = sorted(k for k in sd if k.startswith("decoder."))
    s = sum(float(sd[k].double().sum()) for k in keys)
    a = sum(float(sd[k].double().abs().sum()) for k in keys)
    return np.array([s, a], dtype=np.float64)


def golden_weights(g):
    """Regenerate the fixture's weights from its seed and verify the checksum."""
    sd = random_init_state_dict(SamplerConfig(), seed=int(g["weight_seed"]), head_scale=float(g["head_scale"]),
                                 lattice_identity=bool(int(g["lattice_identity"])) if "lattice_identity" in g else False,
                                 lattice_gamma=float(g["lattice_gamma"]) if "lattice_gamma" in g else 1.0)
    cs = weight_checksum(sd)
    assert np.allclose(cs, g["weight_checksum"], rtol=1e-12, atol=0), \
        f"regenerated weights differ from the fixture's ({cs} vs {g['weight_checksum']})"
    sd["sigma_scheduler.sigmas_norm"] = torch.from_numpy(g["sigmas_norm"]).clone()
    return sd


def forward_golden_cases():
    """(tag, weights, inputs, reference outputs) of tests/golden/ang_forward.npz: reference
    CSPNet.forward on hand-built states with 3..25 A sheared cells."""
    g = load_golden(FORWARD_GOLDEN)
    for tag in [str(c) for c in g["cases"]]:
        sd = random_init_state_dict(SamplerConfig(), seed=int(g[f"{tag}_weight_seed"]))
        assert np.allclose(weight_checksum(sd), g[f"{tag}_weight_checksum"], rtol=1e-12, atol=0)
        yield tag, sd, {k: torch.from_numpy(g[f"{tag}_{k}"]) for k in
                        ("natoms", "a", "x", "l", "t", "text", "types", "lattice", "coords")}


def rel_err(a, b):
    """max|a-b| / max|b| (max-abs error over the max-abs reference value, NOT element-wise
    relative) -- the relative error the parity tolerances are stated in."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
