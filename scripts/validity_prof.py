"""Time the validity pre-filter at benchmark size (4096 crystals x 20 atoms) against the CPU restatement."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chemeleon_b200.validity import validity_flags
from oracle import validity_oracle as VO

B, n = 4096, 20
g = torch.Generator().manual_seed(0)
a = torch.randint(1, 90, (B * n,), generator=g).cuda()
x = torch.rand(B * n, 3, generator=g).cuda()
lat = (torch.eye(3) * 6 + torch.randn(B, 3, 3, generator=g)).cuda()
natoms = [n] * B
for _ in range(3):
    out = validity_flags(a, x, lat, natoms, target="TiO2")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = validity_flags(a, x, lat, natoms, target="TiO2")
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"GPU: {ms*1e3:.1f} us per batch of {B} crystals (incl. host wrapper), {int((out[0]==0).sum())} pass dist/lattice+comp")
t0 = time.perf_counter()
an, xn, ln = a.cpu().numpy(), x.cpu().numpy(), lat.cpu().numpy()
for b in range(256):
    VO.validity_flags(an[b*n:(b+1)*n], xn[b*n:(b+1)*n], ln[b])
dt = time.perf_counter() - t0
print(f"CPU oracle: {dt/256*1e3:.2f} ms per crystal -> {dt/256*B:.1f} s per batch of {B}")
