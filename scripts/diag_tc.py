"""Diagnostic: fp32 vs tc decoder outputs on a golden fixture state."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import load_golden, golden_weights, rel_err
from oracle import chemeleon_oracle as O
from chemeleon_b200.cspnet import CSPNetB200

case, ts = sys.argv[1], [int(v) for v in sys.argv[2:]]
g = load_golden(case); sd = golden_weights(g)
nets = {p: CSPNetB200(sd, precision=p) for p in ("fp32", "tc")}
natoms = g["natoms"].tolist(); nat = torch.tensor(natoms); B = len(natoms)
bi = torch.arange(B).repeat_interleave(nat)
text = torch.from_numpy(g["text"])
for t in ts:
    a, x, l = (torch.from_numpy(g[f"rec{t}_{k}"]) for k in ("a_t", "x_t", "l_t"))
    temb = O.time_embedding(torch.full((B,), t), 128)
    outs = {p: nets[p](a, x, l, nat, bi, t=temb, text_embeds=text) for p in nets}
    print(f"t={t} |l|max={float(l.abs().max()):.2f} a_t nonzero={int((a!=0).sum())}")
    for name, i in (("types", 0), ("lattice", 1), ("coords", 2), ("feat", 3)):
        r, o = outs["fp32"][i].cpu(), outs["tc"][i].cpu()
        print(f"   {name:8s} max|ref|={float(r.abs().max()):.3e} relerr={rel_err(o, r):.3e} nan={int(torch.isnan(o).sum())}")
