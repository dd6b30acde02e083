"""Weight source and packing for the B200 CSPNet decoder.

Accepts the reference checkpoint's `state_dict` key layout (SURVEY.md 8b;
`decoder.csp_layer_{i}.edge_mlp.0.weight [512,1801]`, ...) -- or a live
reference `Chemeleon` / `CSPNet` module -- and produces the device-resident
operands the kernels read:

* the edge-MLP input weight W1 (cspnet.py:113) is split by input block into
  W_hi | W_hj | W_ip | W_fd (columns 0:512 | 512:1024 | 1024:1033 | 1033:1801),
  so that the h_i / h_j / lattice parts become node- and crystal-level terms;
* fp32 row-major copies for the exact (CUDA-core) kernels;
* fp16 copies pre-tiled into the tcgen05 shared-memory operand image
  ("K-major, no swizzle": [K/8][rows][8 halves]) so that one bulk async copy
  brings a whole pipeline stage;
* the FiLM conditioning MLP (cspnet.py:70-73) folded with the sinusoidal time
  embedding into a [T+1, 1024] table (the time half of its input is the same for
  every crystal at a given step, chemeleon.py:380-381).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from .config import SamplerConfig
from . import schedules

Tensor = torch.Tensor


def random_init_state_dict(cfg: SamplerConfig = SamplerConfig(), seed: int = 0, head_scale: float = 1.0,
                           perturb_ln: bool = True, lattice_identity: bool = False,
                           lattice_gamma: float = 1.0, text_tail_dim: int = 0) -> Dict[str, Tensor]:
    """Random weights of the reference architecture, keyed like its checkpoint.

    Checkpoints are not available offline, so benchmarks and parity tests use
    this (north_star: "random-init weights of the same architecture").  Linear
    layers follow torch's default U(-1/sqrt(fan_in), 1/sqrt(fan_in)); the
    embedding is N(0,1).  LayerNorm affine terms are perturbed so tests are
    sensitive to them.  `head_scale` < 1 gives the "tamed" variant whose 1000-step
    dynamics stay bounded (hard parts)."""
    g = torch.Generator().manual_seed(seed)
    H, A = cfg.hidden_dim, cfg.max_atoms
    sd: Dict[str, Tensor] = {}

    def lin(name, out_f, in_f, bias=True, scale=1.0):
        bound = 1.0 / math.sqrt(in_f)
        sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound * scale
        if bias:
            sd[name + ".bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * bound * scale

    def lnorm(name, dim):
        if perturb_ln:
            sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(dim, generator=g)
            sd[name + ".bias"] = 0.1 * torch.randn(dim, generator=g)
        else:
            sd[name + ".weight"] = torch.ones(dim)
            sd[name + ".bias"] = torch.zeros(dim)

    sd["decoder.node_embedding.weight"] = torch.randn(A, H, generator=g)
    cond_in = cfg.time_dim + (cfg.text_dim if cfg.text_guide else 0)
    lin("decoder.film_layer.mlp_cond.0", 2 * H, cond_in)
    lin("decoder.film_layer.proj", H, H)
    lnorm("decoder.film_layer.norm", H)
    for i in range(cfg.num_layers):
        p = f"decoder.csp_layer_{i}"
        lin(p + ".edge_mlp.0", H, cfg.edge_in_dim)
        lin(p + ".edge_mlp.2", H, H)
        lin(p + ".node_mlp.0", H, 2 * H)
        lin(p + ".node_mlp.2", H, H)
        lnorm(p + ".layer_norm", H)
    lnorm("decoder.final_layer_norm", H)
    lin("decoder.coord_out", 3, H, bias=False, scale=head_scale)
    lin("decoder.lattice_out", 9, H, bias=False, scale=head_scale)
    lin("decoder.type_out", A, H, scale=head_scale)
    if lattice_identity:
        # An untrained lattice head predicts ~0 noise, so the ancestral DDPM step
        # (chemeleon.py:420) multiplies the lattice by prod 1/sqrt(alpha_t) ~ 1e3 over the
        # run.  Make the head behave like a trained denoiser for data at the origin:
        # feature 0 of the final LayerNorm is the constant 4 and lattice_out maps it to
        # vec(I)/4, so lattice_out ~= I @ L and the step is a contraction (|l| stays O(1)).
        # `lattice_gamma` < 1 weakens the contraction (lattice_out ~= gamma L): with 0.5 the cosine
        # schedule holds the lattice entries at 5..20 for the WHOLE run, i.e. cells of 8..25 A with
        # shear -- the regime a trained model's raw-Angstrom lattices live in (dataset_utils.py:25).
        sd["decoder.final_layer_norm.weight"][0] = 0.0
        sd["decoder.final_layer_norm.bias"][0] = 4.0
        sd["decoder.lattice_out.weight"][:, 0] = lattice_gamma * torch.eye(3).reshape(9) / 4.0
    if text_tail_dim:
        # TextEncoder.text_emb / null_text_embeds (text_encoder.py:40-46), torch default inits
        E = int(text_tail_dim)
        lin("text_encoder.text_emb.0", E, E)
        lnorm("text_encoder.text_emb.1", E)
        lin("text_encoder.text_emb.3", cfg.text_dim, E)
        sd["text_encoder.null_text_embeds"] = torch.randn(1, E, generator=g)
    sx = schedules.sigma_buffer(cfg.timesteps, cfg.sigma_begin, cfg.sigma_end)
    sd["sigma_scheduler.sigmas"] = sx
    sd["sigma_scheduler.sigmas_norm"] = schedules.sigma_norm_monte_carlo(sx[1:], seed=seed)
    return sd


def state_dict_from(source) -> Dict[str, Tensor]:
    """`source` may be a state_dict, a Lightning checkpoint dict, or a module."""
    if isinstance(source, dict):
        if "state_dict" in source and isinstance(source["state_dict"], dict):
            return source["state_dict"]
        return source
    if hasattr(source, "state_dict"):
        sd = source.state_dict()
        if any(k.startswith("decoder.") for k in sd):
            return sd
        return {"decoder." + k: v for k, v in sd.items()}  # a bare CSPNet
    raise TypeError(f"cannot extract weights from {type(source)}")


def tile_k_major(w: Tensor, dtype=torch.float16) -> Tensor:
    """[rows, K] -> tcgen05 K-major no-swizzle operand image [K/8, rows, 8].

    Core matrices are 8 rows x 16 B; consecutive rows are 16 B apart (SBO = 128 B
    between 8-row groups) and consecutive 8-element K chunks are rows*16 B apart
    (LBO).  A K-chunk of KC columns for all rows is one contiguous block."""
    rows, K = w.shape
    assert K % 8 == 0
    return w.to(dtype).view(rows, K // 8, 8).permute(1, 0, 2).contiguous()


def fd_column_order(num_freqs: int) -> Tensor:
    """Column permutation of W_fd for the tensor-core path.

    Reference order (cspnet.py:49-51): sin block [d*F+k], then cos block
    [3F + d*F+k].  Kernel order: K' = d*2F + 2k + s (s=0 sin, 1 cos) so that one
    thread's rotation recurrence over k emits contiguous 16-byte pieces."""
    F = num_freqs
    idx = torch.empty(6 * F, dtype=torch.long)
    for d in range(3):
        for k in range(F):
            idx[d * 2 * F + 2 * k + 0] = d * F + k
            idx[d * 2 * F + 2 * k + 1] = 3 * F + d * F + k
    return idx


@dataclass
class LayerWeights:
    # fp32 row-major [out, in]
    w_hij: Tensor      # [1024,512]  rows 0:512 -> P_i (h_i block), 512:1024 -> P_j
    w_ip: Tensor       # [512,9]
    b1: Tensor         # [512]
    w_fd: Tensor       # [512,768]   reference column order
    w2: Tensor
    b2: Tensor
    wn1: Tensor        # [512,1024]
    bn1: Tensor
    wn2: Tensor
    bn2: Tensor
    ln_g: Tensor
    ln_b: Tensor
    # fp16 tcgen05 operand images
    w_hij_t: Optional[Tensor] = None
    w_fd_t: Optional[Tensor] = None   # kernel column order (fd_column_order)
    w2_t: Optional[Tensor] = None
    wn1_t: Optional[Tensor] = None
    wn2_t: Optional[Tensor] = None


@dataclass
class TextTailWeights:
    """`TextEncoder.text_emb` + `null_text_embeds` (text_encoder/text_encoder.py:40-46): everything of the
    text conditioning that comes after the language model."""
    embed_dim: int
    text_dim: int
    w1: Tensor
    b1: Tensor
    ln_g: Tensor
    ln_b: Tensor
    w2: Tensor
    b2: Tensor
    null_embeds: Tensor            # [1, embed_dim]


@dataclass
class PackedWeights:
    cfg: SamplerConfig
    device: torch.device
    emb: Tensor                    # [104,512]
    film_wp: Tensor                # [512,512]
    film_bp: Tensor
    film_g: Tensor
    film_b: Tensor
    film_wp_t: Optional[Tensor]
    film_time_table: Tensor        # [T+1,1024] = W_cond[:, :time_dim] @ time_emb(t)
    film_w_text: Optional[Tensor]  # [1024,text_dim]
    film_b_cond: Tensor            # [1024]
    layers: List[LayerWeights]
    final_g: Tensor
    final_b: Tensor
    w_head: Tensor                 # [128,512]: rows 0:104 type_out, 104:107 coord_out, rest 0
    b_head: Tensor                 # [128]
    w_head_t: Optional[Tensor]
    w_lat: Tensor                  # [9,512]
    sigmas_norm: Tensor            # [T+1] (host)
    q_mats: Optional[Tensor] = None
    q_one_step_mats: Optional[Tensor] = None
    extra: Dict[str, Tensor] = field(default_factory=dict)
    text_tail: Optional[TextTailWeights] = None   # present when the state_dict holds text_encoder.text_emb.*


def pack_weights(source, cfg: Optional[SamplerConfig] = None, device="cuda", tensor_core: bool = True) -> PackedWeights:
    """Split / fold / tile the reference weights and move them to `device`."""
    cfg = cfg or SamplerConfig()
    cfg.validate()
    sd = state_dict_from(source)
    dev = torch.device(device)
    H, F = cfg.hidden_dim, cfg.num_freqs

    def get(name) -> Tensor:
        if name not in sd:
            raise KeyError(f"weight {name!r} missing from state_dict")
        return sd[name].detach().to(torch.float32).cpu()

    def d(t: Tensor) -> Tensor:
        return t.contiguous().to(dev)

    def tiled(t: Tensor) -> Optional[Tensor]:
        return d(tile_k_major(t)) if tensor_core else None

    def tiled_blocks(t: Tensor, rows: int = 128) -> Optional[Tensor]:
        """One K-major image per block of `rows` output channels (the edge kernel streams
        W2 one 128-channel output unit at a time): [out/rows][K/8][rows][8]."""
        if not tensor_core:
            return None
        return d(torch.cat([tile_k_major(t[i:i + rows]) for i in range(0, t.shape[0], rows)], dim=0))

    def head_split_image(w: Tensor) -> Optional[Tensor]:
        """Operand blob of the split-precision head GEMM: 16-byte header (float32 1/s) followed by
        the K-major image of s*[w_hi | w_hi | w_lo] padded to 256 rows ([192][256][8] fp16).  The
        power-of-two scale s keeps w_lo = s*w - w_hi out of the fp16 subnormals."""
        if not tensor_core:
            return None
        wmax = float(w.abs().max())
        s = 2.0 ** math.floor(math.log2(1024.0 / wmax)) if wmax > 0 else 1.0
        ws = w.double() * s
        hi = ws.to(torch.float16)
        lo = (ws - hi.double()).to(torch.float16)
        full = torch.zeros(256, 3 * w.shape[1], dtype=torch.float16)
        full[: w.shape[0]] = torch.cat([hi, hi, lo], dim=1)
        blob = torch.zeros(8 + full.numel(), dtype=torch.float16)
        blob[:2].view(torch.float32)[0] = 1.0 / s
        blob[8:] = tile_k_major(full).reshape(-1)
        return d(blob)

    w_cond = get("decoder.film_layer.mlp_cond.0.weight")  # [1024, time_dim(+text_dim)]
    b_cond = get("decoder.film_layer.mlp_cond.0.bias")
    te = schedules.time_embedding_table(cfg.timesteps, cfg.time_dim)  # [T+1,128]
    time_table = te @ w_cond[:, : cfg.time_dim].t()
    w_text = w_cond[:, cfg.time_dim:] if w_cond.shape[1] > cfg.time_dim else None

    perm = fd_column_order(F)
    layers = []
    for i in range(cfg.num_layers):
        p = f"decoder.csp_layer_{i}"
        w1 = get(p + ".edge_mlp.0.weight")
        assert w1.shape == (H, cfg.edge_in_dim), w1.shape
        w_hi, w_hj = w1[:, :H], w1[:, H:2 * H]
        w_ip = w1[:, 2 * H:2 * H + 9]
        w_fd = w1[:, 2 * H + 9:]
        w_hij = torch.cat([w_hi, w_hj], dim=0)
        w2 = get(p + ".edge_mlp.2.weight")
        wn1 = get(p + ".node_mlp.0.weight")
        wn2 = get(p + ".node_mlp.2.weight")
        layers.append(LayerWeights(
            w_hij=d(w_hij), w_ip=d(w_ip), b1=d(get(p + ".edge_mlp.0.bias")), w_fd=d(w_fd),
            w2=d(w2), b2=d(get(p + ".edge_mlp.2.bias")),
            wn1=d(wn1), bn1=d(get(p + ".node_mlp.0.bias")),
            wn2=d(wn2), bn2=d(get(p + ".node_mlp.2.bias")),
            ln_g=d(get(p + ".layer_norm.weight")), ln_b=d(get(p + ".layer_norm.bias")),
            w_hij_t=tiled(w_hij), w_fd_t=tiled(w_fd[:, perm]),
            w2_t=tiled_blocks(w2 * 0.5),      # image of W2 / 2 (exact in fp16): the GEMM2 accumulator is the x / 2 of SiLU's tanh form

            wn1_t=tiled(wn1), wn2_t=tiled(wn2)))

    A = cfg.max_atoms
    w_head = torch.zeros(128, H)
    b_head = torch.zeros(128)
    w_head[:A] = get("decoder.type_out.weight")
    b_head[:A] = get("decoder.type_out.bias")
    w_head[A:A + 3] = get("decoder.coord_out.weight")

    if "sigma_scheduler.sigmas_norm" in sd:
        sn = sd["sigma_scheduler.sigmas_norm"].detach().to(torch.float32).cpu()
    else:
        sn = schedules.sigma_norm_monte_carlo(
            schedules.sigma_buffer(cfg.timesteps, cfg.sigma_begin, cfg.sigma_end)[1:])

    tail = None
    if "text_encoder.text_emb.0.weight" in sd and "text_encoder.null_text_embeds" in sd:
        tw1 = get("text_encoder.text_emb.0.weight")
        tw2 = get("text_encoder.text_emb.3.weight")
        tail = TextTailWeights(
            embed_dim=int(tw1.shape[1]), text_dim=int(tw2.shape[0]), w1=d(tw1), b1=d(get("text_encoder.text_emb.0.bias")),
            ln_g=d(get("text_encoder.text_emb.1.weight")), ln_b=d(get("text_encoder.text_emb.1.bias")),
            w2=d(tw2), b2=d(get("text_encoder.text_emb.3.bias")),
            null_embeds=d(get("text_encoder.null_text_embeds").reshape(1, -1)))

    wp = get("decoder.film_layer.proj.weight")
    return PackedWeights(
        cfg=cfg, device=dev, emb=d(get("decoder.node_embedding.weight")),
        film_wp=d(wp), film_bp=d(get("decoder.film_layer.proj.bias")),
        film_g=d(get("decoder.film_layer.norm.weight")), film_b=d(get("decoder.film_layer.norm.bias")),
        film_wp_t=tiled(wp), film_time_table=d(time_table),
        film_w_text=d(w_text) if w_text is not None else None, film_b_cond=d(b_cond),
        layers=layers, final_g=d(get("decoder.final_layer_norm.weight")),
        final_b=d(get("decoder.final_layer_norm.bias")),
        w_head=d(w_head), b_head=d(b_head), w_head_t=head_split_image(w_head),
        w_lat=d(get("decoder.lattice_out.weight")), sigmas_norm=sn,
        q_mats=sd.get("d3pm.q_mats"), q_one_step_mats=sd.get("d3pm.q_one_step_mats"),
        extra={"film_w_cond": d(w_cond)}, text_tail=tail)
