"""Development aid (CPU): emulate the rounding points of the tensor-core decoder path in PyTorch
and report which of them dominates the error against the fp32 oracle for a given golden record.

    python scripts/tc_numerics.py c1_full_6 995
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import math
import torch
import torch.nn.functional as F
from helpers import golden_weights, load_golden, rel_err
from oracle import chemeleon_oracle as O

def r16(x, on=True):
    return x.half().float() if on else x

def silu(x):
    return x * torch.sigmoid(x)

def forward(w, a, x, l, nat, bi, temb, text, opt):
    """opt: dict of switches; True = round like the kernel does."""
    H = 512
    N, B = a.shape[0], l.shape[0]
    edges = O.fc_edges(nat)
    ei, ej = edges
    fd = (x[ej] - x[ei])
    emb = O.sinusoid_embedding(fd % 1.0, 128)
    emb = r16(emb, opt["emb"])
    h = w["decoder.node_embedding.weight"][a]
    cond_in = torch.cat([temb, text], 1)
    cond = silu(F.linear(cond_in, w["decoder.film_layer.mlp_cond.0.weight"], w["decoder.film_layer.mlp_cond.0.bias"]))
    scale, shift = cond[bi].chunk(2, 1)
    ip = (l @ l.transpose(1, 2)).reshape(B, 9)
    for i in range(6):
        p = f"decoder.csp_layer_{i}"
        y = F.linear(r16(h, opt["h16"]), r16(w["decoder.film_layer.proj.weight"], opt["w"]), w["decoder.film_layer.proj.bias"])
        y = F.layer_norm(y, (H,), w["decoder.film_layer.norm.weight"], w["decoder.film_layer.norm.bias"])
        h = silu(y * scale + shift) + h
        hn = F.layer_norm(h, (H,), w[p + ".layer_norm.weight"], w[p + ".layer_norm.bias"])
        hn16 = r16(hn, opt["hn"])
        W1 = w[p + ".edge_mlp.0.weight"]
        Whi, Whj, Wip, Wfd = W1[:, :512], W1[:, 512:1024], W1[:, 1024:1033], W1[:, 1033:]
        Pi = r16(F.linear(hn16, r16(Whi, opt["w"])), opt["P"])
        Pj = r16(F.linear(hn16, r16(Whj, opt["w"])), opt["P"])
        cg = F.linear(ip, Wip, w[p + ".edge_mlp.0.bias"])
        U = F.linear(emb, r16(Wfd, opt["w"]))
        pre = U + Pi[ei] + cg[bi[ei]] + Pj[ej]
        if opt["pre16"]:
            pre = r16(pre)
        a1 = r16(silu(pre), opt["a1"])
        if opt.get("a1_split"):
            a1 = silu(pre)
            hi = a1.half().float(); lo = (a1 - hi).half().float(); a1 = hi + lo
        e = silu(F.linear(a1, r16(w[p + ".edge_mlp.2.weight"], opt["w"]), w[p + ".edge_mlp.2.bias"]))
        agg = O.scatter_mean(e, ei, N)
        agg = r16(agg, opt["agg"])
        z = silu(F.linear(torch.cat([hn16, agg], 1), r16(w[p + ".node_mlp.0.weight"], opt["w"]), w[p + ".node_mlp.0.bias"]))
        z = r16(z, opt["z"])
        h = h + silu(F.linear(z, r16(w[p + ".node_mlp.2.weight"], opt["w"]), w[p + ".node_mlp.2.bias"]))
    hf = F.layer_norm(h, (H,), w["decoder.final_layer_norm.weight"], w["decoder.final_layer_norm.bias"])
    coords = F.linear(hf, w["decoder.coord_out.weight"])
    gf = O.scatter_mean(hf, bi, B)
    lat = F.linear(gf, w["decoder.lattice_out.weight"]).view(-1, 3, 3) @ l
    types = F.linear(hf, w["decoder.type_out.weight"], w["decoder.type_out.bias"])
    return types, lat, coords

ALL = ["emb", "h16", "w", "hn", "P", "pre16", "a1", "agg", "z"]

def main():
    case, t = sys.argv[1], int(sys.argv[2])
    g = load_golden(case)
    sd = golden_weights(g)
    w = {k: v for k, v in sd.items() if k.startswith("decoder.")}
    natoms = g["natoms"].tolist()
    nat = torch.tensor(natoms); B = len(natoms)
    bi = torch.arange(B).repeat_interleave(nat)
    a, x, l = (torch.from_numpy(g[f"rec{t}_{k}"]) for k in ("a_t", "x_t", "l_t"))
    temb = O.time_embedding(torch.full((B,), t), 128)
    text = torch.from_numpy(g["text"])
    print("lattice max", float(l.abs().max()))
    ref = forward(w, a, x, l, nat, bi, temb, text, {k: False for k in ALL})
    def err(o):
        return [rel_err(o[i], ref[i]) for i in range(3)]
    print("all rounding   ", ["%.2e" % v for v in err(forward(w, a, x, l, nat, bi, temb, text, {k: True for k in ALL}))])
    for k in ALL:
        only = {kk: kk == k for kk in ALL}
        print(f"only {k:6s}    ", ["%.2e" % v for v in err(forward(w, a, x, l, nat, bi, temb, text, only))])
    for k in ALL:
        allbut = {kk: kk != k for kk in ALL}
        print(f"all but {k:6s} ", ["%.2e" % v for v in err(forward(w, a, x, l, nat, bi, temb, text, allbut))])
    o = {k: True for k in ALL}; o["a1_split"] = True; o["pre16"] = False
    print("a1 split       ", ["%.2e" % v for v in err(forward(w, a, x, l, nat, bi, temb, text, o))])

main()
