"""Host-side owner of the packed weights, the cb2_model struct and the workspace.

Thin: all compute is in libchemeleon_b200.so.  PyTorch is used only for device
memory, streams and (off the hot path) the one-per-prompt text projection.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib
from .config import SamplerConfig
from .topology import BatchTopology
from .weights import PackedWeights, pack_weights


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class DecoderEngine:
    """CSPNet decoder on one B200.  `precision`: "fp32" (exact) or "tc" (tcgen05 fp16)."""

    def __init__(self, source, cfg: Optional[SamplerConfig] = None, device="cuda", precision: str = "fp32"):
        if not torch.cuda.is_available():
            raise _lib.Cb2Error("chemeleon_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        _lib.check(self.lib.cb2_check_device(self.device.index), "cb2_check_device")
        self.cfg = cfg or SamplerConfig()
        self.precision = {"fp32": _lib.PRECISION_FP32, "tc": _lib.PRECISION_TC_F16}[precision]
        self.precision_name = precision
        with torch.cuda.device(self.device):
            self.w: PackedWeights = source if isinstance(source, PackedWeights) else pack_weights(
                source, self.cfg, self.device, tensor_core=(precision == "tc"))
        self.model = self._make_model()
        self._topo_cache: Dict[Tuple, BatchTopology] = {}
        self._ws: Optional[torch.Tensor] = None

    # -- structs ---------------------------------------------------------------
    def _make_model(self) -> _lib.Model:
        w, cfg = self.w, self.cfg
        m = _lib.Model()
        m.abi_version = _lib.ABI_VERSION
        m.hidden, m.n_layers = cfg.hidden_dim, cfg.num_layers
        m.n_atom_types, m.n_freqs, m.timesteps = cfg.max_atoms, cfg.num_freqs, cfg.timesteps
        p = _lib.ptr
        m.emb = p(w.emb)
        m.film_wp, m.film_bp, m.film_g, m.film_b = p(w.film_wp), p(w.film_bp), p(w.film_g), p(w.film_b)
        m.film_wp_t = p(w.film_wp_t)
        m.film_time_table = p(w.film_time_table)
        for i, L in enumerate(w.layers):
            l = m.layers[i]
            for name in ("w_hij", "w_ip", "b1", "w_fd", "w2", "b2", "wn1", "bn1", "wn2", "bn2", "ln_g", "ln_b",
                         "w_hij_t", "w_fd_t", "w2_t", "wn1_t", "wn2_t"):
                setattr(l, name, p(getattr(L, name)))
        m.final_g, m.final_b = p(w.final_g), p(w.final_b)
        m.w_head, m.b_head, m.w_head_t, m.w_lat = p(w.w_head), p(w.b_head), p(w.w_head_t), p(w.w_lat)
        return m

    def topology(self, natoms: Sequence[int], n_variants: int) -> BatchTopology:
        key = (tuple(int(n) for n in natoms), int(n_variants))
        t = self._topo_cache.get(key)
        if t is None:
            if len(self._topo_cache) > 8:
                self._topo_cache.clear()
            t = BatchTopology(key[0], n_variants, self.device,
                              exact=(self.precision == _lib.PRECISION_FP32),
                              tensor_core=(self.precision == _lib.PRECISION_TC_F16))
            if self.precision == _lib.PRECISION_TC_F16 and t.n_tiles == 0 and t.N > 0:
                raise _lib.Cb2Error("tensor-core path supports crystals of at most 128 atoms")
            self._topo_cache[key] = t
        return t

    def workspace(self, topo: BatchTopology) -> torch.Tensor:
        need = int(self.lib.cb2_workspace_bytes(C.byref(self.model), topo.byref(), self.precision))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    # -- FiLM conditioning -------------------------------------------------------
    def text_part(self, text_embeds: torch.Tensor) -> torch.Tensor:
        """W_cond[:, time_dim:] @ text + b_cond  -> [rows,1024] from projected text embeddings [rows,512]
        (the output of `TextEncoder.get_text_embeds`).  Once per prompt, off the hot path."""
        w = self.w
        if w.film_w_text is None:
            raise ValueError("model has no text conditioning (text_guide=False)")
        x = text_embeds.to(self.device, torch.float32).contiguous()
        out = torch.empty(x.shape[0], 2 * self.cfg.hidden_dim, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cb2_linear_f32(x.data_ptr(), x.shape[1], w.film_w_text.data_ptr(), w.film_b_cond.data_ptr(),
                                               out.data_ptr(), out.shape[1], x.shape[0], out.shape[1], x.shape[1], 0,
                                               _stream_ptr()), "cb2_linear_f32")
        return out

    @property
    def has_text_tail(self) -> bool:
        return self.w.text_tail is not None and self.w.film_w_text is not None

    def text_condition(self, encoder_embeds: torch.Tensor) -> torch.Tensor:
        """Language-model embeddings [P, embed_dim] of P distinct prompts -> FiLM text rows [P + 1, 1024]
        (row P = the unconditional row from the learned null embedding): `TextEncoder.text_emb` +
        `null_text_embeds` + the text half of `FilmLayer.mlp_cond`, on the device (cb2_text_condition)."""
        if not self.has_text_tail:
            raise ValueError("the weights hold no text_encoder.text_emb.* / null_text_embeds (or text_guide=False)")
        t = self.w.text_tail
        enc = encoder_embeds.to(self.device, torch.float32).contiguous()
        if enc.dim() != 2 or enc.shape[1] != t.embed_dim:
            raise ValueError(f"encoder embeddings must be [n_prompts, {t.embed_dim}]")
        tt = _lib.TextTail()
        tt.embed_dim, tt.text_dim = t.embed_dim, t.text_dim
        for name in ("w1", "b1", "ln_g", "ln_b", "w2", "b2", "null_embeds"):
            setattr(tt, name, getattr(t, name).data_ptr())
        tt.w_text, tt.b_cond = self.w.film_w_text.data_ptr(), self.w.film_b_cond.data_ptr()
        P = enc.shape[0]
        out = torch.empty(P + 1, 2 * self.cfg.hidden_dim, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            ws = torch.empty(int(self.lib.cb2_text_condition_workspace_bytes(C.byref(tt), P)), dtype=torch.uint8,
                             device=self.device)
            _lib.check(self.lib.cb2_text_condition(C.byref(tt), enc.data_ptr() if P else None, P, out.data_ptr(),
                                                   ws.data_ptr(), ws.numel(), _stream_ptr()), "cb2_text_condition")
        return out

    def film_cond_from_embeddings(self, t_emb: Optional[torch.Tensor], text: Optional[torch.Tensor],
                                  topo: BatchTopology) -> Optional[torch.Tensor]:
        """SiLU(mlp_cond(cat[t, text])) for arbitrary per-crystal embeddings (module interface)."""
        if t_emb is None and text is None:
            return None
        parts = [x.to(self.device, torch.float32) for x in (t_emb, text) if x is not None]
        cond_in = torch.cat(parts, dim=1).contiguous()
        wc = self.w.extra["film_w_cond"]
        if cond_in.shape[1] != wc.shape[1]:
            raise ValueError(f"conditioning width {cond_in.shape[1]} does not match mlp_cond ({wc.shape[1]})")
        rows = cond_in.shape[0]
        out = torch.empty(rows, 2 * self.cfg.hidden_dim, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.cb2_linear_f32(cond_in.data_ptr(), cond_in.shape[1], wc.data_ptr(),
                                           self.w.film_b_cond.data_ptr(), out.data_ptr(), out.shape[1],
                                           rows, out.shape[1], cond_in.shape[1], 1, _stream_ptr()),
                   "cb2_linear_f32")
        return out

    # -- one forward ---------------------------------------------------------------
    def forward(self, topo: BatchTopology, atom_types: torch.Tensor, frac_coords: torch.Tensor,
                lattices: torch.Tensor, film_cond: Optional[torch.Tensor], coords_only: bool = False,
                want_features: bool = True):
        V, N, B = topo.V, topo.N, topo.B
        dev = self.device
        head = torch.empty(V * N, _lib.HEAD_COLS, device=dev, dtype=torch.float32)
        lat = torch.empty(V * B, 9, device=dev, dtype=torch.float32)
        feat = torch.empty(V * N, self.cfg.hidden_dim, device=dev, dtype=torch.float32) if want_features else None
        ws = self.workspace(topo)
        io = _lib.ForwardIO()
        io.atom_types = atom_types.data_ptr()
        io.frac_coords = frac_coords.data_ptr()
        io.lattices = lattices.data_ptr()
        io.film_cond = _lib.ptr(film_cond)
        io.head_out, io.lattice_out, io.node_features = head.data_ptr(), lat.data_ptr(), _lib.ptr(feat)
        io.coords_only = int(coords_only)
        io.precision = self.precision
        # guard bits per crystal (CB2_FLAG_TC_RANGE: cell outside the fp16 range of the tensor-core path)
        self.last_flags = torch.zeros(max(B, 1), device=dev, dtype=torch.int32)
        io.flags = self.last_flags.data_ptr()
        _lib.check(self.lib.cb2_decoder_forward(C.byref(self.model), topo.byref(), C.byref(io), ws.data_ptr(),
                                                ws.numel(), _stream_ptr()), "cb2_decoder_forward")
        return head, lat, feat
