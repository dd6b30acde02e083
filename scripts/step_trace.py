"""In-step kernel durations of the captured sampler timestep (CUPTI through torch.profiler: the graph replays at
the clocks of the steady state, unlike ncu's serialised cold-cache replays).

    python scripts/step_trace.py [batch] [natoms] [steps]
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from chemeleon_b200 import _lib
if os.environ.get("CB2_LIB"):
    _lib.LIB_PATH = os.environ["CB2_LIB"]      # development: a variant build of the library
from chemeleon_b200 import dist as cdist
from chemeleon_b200.config import SamplerConfig
from chemeleon_b200.sampler import ChemeleonB200
from chemeleon_b200.weights import random_init_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
K = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cfg = SamplerConfig()
sd = random_init_state_dict(cfg, seed=0, head_scale=0.01, lattice_identity=True)
model = ChemeleonB200(sd, cfg, device="cuda:0", precision="tc")
g = torch.Generator().manual_seed(1)
text = torch.randn(B, cfg.text_dim, generator=g).pin_memory()
null = torch.randn(1, cfg.text_dim, generator=g).pin_memory()
plan = cdist.ShardPlan([n] * B, 1, 0)
run = cdist.prepare_sharded_run(model, plan, text, null, 2.0, 1e-5, seed=1234)
run.capture()
for _ in range(6):
    run.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    run.step()
e1.record()
torch.cuda.synchronize()
print(f"untraced: {e0.elapsed_time(e1) / K:.3f} ms per timestep")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(K):
        run.step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
tot = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    k = e.name.split("(")[0].replace("cb2::", "")
    tot[k][0] += 1
    tot[k][1] += e.time_range.end - e.time_range.start
span = (ev[-1].time_range.end - ev[0].time_range.start) / K
busy = sum(v[1] for v in tot.values()) / K
print(f"traced: span {span / 1e3:.3f} ms per timestep, kernels busy {busy / 1e3:.3f} ms")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} n/step={c / K:5.1f} avg={t / c:9.1f} us  per step={t / K / 1e3:7.3f} ms  {t / K / span:.3f}")
first = [e for e in ev if "k_film_cond" in e.name]
if len(first) >= 2:
    a, b = first[0].time_range.start, first[1].time_range.start
    print("one timestep, in order:")
    for e in ev:
        if a <= e.time_range.start < b:
            print(f"  {e.name.split('(')[0].replace('cb2::', ''):32s} {e.time_range.end - e.time_range.start:9.1f} us")
