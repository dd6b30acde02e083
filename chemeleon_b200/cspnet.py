"""`CSPNetB200`: drop-in for the reference decoder's module interface.

Mirrors `CSPNet.forward(atom_types, frac_coords, lattices, num_atoms, node2graph,
t=None, text_embeds=None) -> DECODER_OUTPUTS` (chemeleon/modules/cspnet.py:345-405,
16-18) so it can be assigned to `model.decoder` of a reference `Chemeleon`.
Inference only: no autograd graph is built, and inputs that require grad are
rejected.  Edges are implied by `num_atoms` (edge_style="fc"); the reference's
`gen_edges` / `edge_index` machinery has no counterpart here.
"""
from __future__ import annotations

from collections import namedtuple
from typing import Optional

import torch
import torch.nn as nn

from .config import SamplerConfig
from .engine import DecoderEngine

DECODER_OUTPUTS = namedtuple(
    "DECODER_OUTPUTS", ["atom_types_out", "lattice_out", "coords_out", "node_features"]
)


class CSPNetB200(nn.Module):
    def __init__(self, source, cfg: Optional[SamplerConfig] = None, device="cuda", precision: str = "fp32"):
        """`source`: reference `CSPNet`/`Chemeleon` module, its state_dict, or a checkpoint dict."""
        super().__init__()
        self.engine = DecoderEngine(source, cfg, device, precision)
        self.cfg = self.engine.cfg
        self.edge_style = "fc"
        self.num_layers = self.cfg.num_layers

    @torch.no_grad()
    def forward(self, atom_types, frac_coords, lattices, num_atoms, node2graph=None, t=None, text_embeds=None):
        for x in (frac_coords, lattices, t, text_embeds):
            if x is not None and x.requires_grad:
                raise RuntimeError("CSPNetB200 is inference-only (no backward kernels)")
        if atom_types.dim() != 1:
            raise ValueError("smooth=False decoder expects integer atom types [N]")
        eng = self.engine
        dev = eng.device
        natoms = [int(n) for n in num_atoms.tolist()]
        topo = eng.topology(natoms, 1)
        if node2graph is not None and node2graph.numel() != topo.N:
            raise ValueError("node2graph does not match num_atoms")
        a = atom_types.to(dev, torch.int64).contiguous()
        x = frac_coords.to(dev, torch.float32).contiguous()
        l = lattices.to(dev, torch.float32).reshape(-1, 9).contiguous()
        with torch.cuda.device(dev):
            cond = eng.film_cond_from_embeddings(t, text_embeds, topo)
            head, lat, feat = eng.forward(topo, a, x, l, cond, coords_only=False, want_features=True)
        A = self.cfg.max_atoms
        type_out = head[:, :A] if self.cfg.pred_atom_types else None
        return DECODER_OUTPUTS(atom_types_out=type_out, lattice_out=lat.view(-1, 3, 3),
                               coords_out=head[:, A:A + 3], node_features=feat)
