"""The metric measured literally: ONE complete 1000-step sampling job of BASELINE config 3 (4096 x 20-atom cells)
through the public API (host embeddings in, list of Atoms out), timed end to end.
    python scripts/full_job.py [batch] [natoms]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chemeleon_b200.config import SamplerConfig
from chemeleon_b200.sampler import ChemeleonB200
from chemeleon_b200.weights import random_init_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg = SamplerConfig()
model = ChemeleonB200(random_init_state_dict(cfg, seed=0, head_scale=0.01, lattice_identity=True), cfg, precision="tc")
g = torch.Generator().manual_seed(1)
text = torch.randn(B, cfg.text_dim, generator=g).pin_memory()
null = torch.randn(1, cfg.text_dim, generator=g).pin_memory()
model.sample_batch([n] * B, text_embeds=text, null_text_embeds=null, seed=1, t_stop=cfg.timesteps - 3)   # warm: graph capture
torch.cuda.synchronize()
t0 = time.perf_counter()
atoms, flags = model.sample_batch([n] * B, text_embeds=text, null_text_embeds=null, seed=2, return_flags=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{B} x {n}-atom cells, {cfg.timesteps} timesteps, host embeddings -> {len(atoms)} Atoms: {dt:.2f} s = "
      f"{B / dt:.2f} structures/s; flagged crystals: {int((flags != 0).sum())}")
