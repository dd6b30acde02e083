"""Timesteps per second with and without per-step streaming (frames packed on the device, async copy
into a pinned ring on a side stream) -- SURVEY.md 8f row 4.   python scripts/stream_bench.py [B] [n] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chemeleon_b200.config import SamplerConfig
from chemeleon_b200.sampler import ChemeleonB200
from chemeleon_b200.weights import random_init_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
cfg = SamplerConfig(timesteps=steps)
model = ChemeleonB200(random_init_state_dict(cfg, seed=0, head_scale=0.01, lattice_identity=True), cfg, precision="tc")
g = torch.Generator().manual_seed(1)
text, null = torch.randn(B, 512, generator=g), torch.randn(1, 512, generator=g)
natoms = [n] * B
model.sample_states(natoms, text, null, seed=1)                       # warm: capture the graph
torch.cuda.synchronize(); t0 = time.perf_counter()
model.sample_states(natoms, text, null, seed=2)
torch.cuda.synchronize(); t_plain = (time.perf_counter() - t0) / steps
t0 = time.perf_counter(); nb = 0
for f in model._sample_generator(natoms, None, 2.0, 1e-5, text_embeds=text, null_text_embeds=null, seed=2, frames=True):
    nb += len(f.buf)
torch.cuda.synchronize(); t_frames = (time.perf_counter() - t0) / steps
t0 = time.perf_counter()
for ats in model._sample_generator(natoms, None, 2.0, 1e-5, text_embeds=text, null_text_embeds=null, seed=2):
    pass
torch.cuda.synchronize(); t_atoms = (time.perf_counter() - t0) / steps
# the reference's way: blocking copy of the state after every step
run = model.make_run(natoms, text, null, 2.0, 1e-5, None, 3)
run.init_state(*model.initial_noise(run.B, run.N, 3))
t0 = time.perf_counter()
for _ in range(steps):
    run.step(); a, x, l = run.a.cpu(), run.x.cpu(), run.l.cpu()
t_block = (time.perf_counter() - t0) / steps
print(f"B={B} n={n}: {t_plain*1e3:.2f} ms/step no streaming | {t_frames*1e3:.2f} frames ({nb/steps/1e6:.2f} MB/step) | "
      f"{t_atoms*1e3:.2f} frames + Atoms objects | {t_block*1e3:.2f} blocking .cpu() per step")
