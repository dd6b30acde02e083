"""BASELINE config 2 (composition TiO2, 100 samples per Z-factor bucket, n = 3..39): ONE ragged batch
(driver.sample_compositions) against the reference's way of 13 sequential sample() calls
(sample_target_composition.py:37-51), same sampler underneath.   python scripts/c2_compare.py [timesteps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chemeleon_b200 import driver
from chemeleon_b200.config import SamplerConfig
from chemeleon_b200.sampler import ChemeleonB200
from chemeleon_b200.weights import random_init_state_dict

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cfg = SamplerConfig(timesteps=steps)
sd = random_init_state_dict(cfg, seed=0, head_scale=0.01, lattice_identity=True, text_tail_dim=768)
model = ChemeleonB200(sd, cfg, precision="tc")
model.set_prompt_embedding("O2 Ti1", torch.randn(768, generator=torch.Generator().manual_seed(0)))
for label in ("warm (graph capture included)", "steady"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = driver.sample_compositions(model, ["TiO2"], n_samples=100, max_natoms=40, max_factor=13, seed=1)
    torch.cuda.synchronize(); t_one = time.perf_counter() - t0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n_seq = 0
    for f in range(1, 14):
        atoms = model.sample("O2 Ti1", 3 * f, 100, seed=f)
        n_seq += len(atoms)
    torch.cuda.synchronize(); t_seq = time.perf_counter() - t0
    print(f"[{label}] {steps} timesteps: one ragged batch of {len(res['natoms'])} structures {t_one:.3f} s | "
          f"13 sequential sample() calls ({n_seq} structures) {t_seq:.3f} s | ratio {t_seq / t_one:.2f}x")
