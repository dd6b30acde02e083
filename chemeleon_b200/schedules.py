"""Noise-schedule tables and the per-timestep coefficient table the update kernel reads.

Host-side, computed once per sampler (off the hot path) with fp32 torch ops in
the same operation order as the reference so that the scalars agree to the bit:

* BetaScheduler buffers      -- chemeleon/utils/diff_utils.py:10-19, 57-102
* SigmaScheduler.sigmas      -- chemeleon/utils/diff_utils.py:109-127
* lattice / coordinate step scalars -- chemeleon/modules/chemeleon.py:413-456
* D3PM absorbing-state matrices in closed form -- diff_utils.py:195-213, 168-185:
  Q_s = (1-b_s) I + b_s 1 e0^T  and  Qbar_s = Q_1...Q_s = diag_s I + off_s 1 e0^T
  (row 0: Qbar[0,0]); the recurrence below is the matmul chain restricted to the
  three distinct entries, so no [1001,104,104] table is ever built.

`sigmas_norm` is a Monte-Carlo buffer in the reference (diff_utils.py:49-54,119)
and is therefore taken from the checkpoint / module whenever one is given.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

# column indices of the coefficient table (one row per timestep t = 0..T)
C_LAT_C0 = 0       # 1/sqrt(alpha_t)
C_LAT_C1 = 1       # (1-alpha_t)/sqrt(1-alphabar_t)
C_LAT_SIG = 2      # sqrt(beta_t (1-alphabar_{t-1})/(1-alphabar_t))
C_X_STEP = 3       # sigma_t^2 - sigma_{t-1}^2
C_X_STD = 4        # sqrt(sigma_{t-1}^2 (sigma_t^2 - sigma_{t-1}^2)/sigma_t^2)
C_X_SQRT_NORM = 5  # sqrt(sigmas_norm[t])
C_X_STEP2 = 6      # step_lr (sigma_t/sigma_begin)^2
C_X_STD2 = 7       # sqrt(2 step2)
C_D_BETA = 8       # Q_{t-1}[r,0], r != 0          (= beta_{t-1})
C_D_OMB = 9        # Q_{t-1}[c,c], c != 0          (= 1-beta_{t-1})
C_D_Q00 = 10       # Q_{t-1}[0,0]
C_D_DIAG = 11      # Qbar_{t-2}[c,c], c != 0      (= alphabar_{t-2})
C_D_OFF = 12       # Qbar_{t-2}[c,0], c != 0      (= 1-alphabar_{t-2})
C_D_QB00 = 13      # Qbar_{t-2}[0,0]
C_NCOLS = 16


def cosine_beta_schedule(timesteps: int, s: float = 0.008) -> torch.Tensor:
    x = torch.linspace(0, timesteps, timesteps + 1)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return torch.clip(betas, 0.0001, 0.9999)


def beta_buffers(timesteps: int, mode: str = "cosine") -> Dict[str, torch.Tensor]:
    if mode != "cosine":
        raise ValueError(f"beta_schedule={mode!r} is not supported by the B200 sampler")
    betas = torch.cat([torch.zeros(1), cosine_beta_schedule(timesteps)], dim=0)
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    sig = torch.zeros_like(betas)
    sig[1:] = betas[1:] * (1.0 - ac[:-1]) / (1.0 - ac[1:])
    return dict(betas=betas, alphas=alphas, alphas_cumprod=ac, sigmas=torch.sqrt(sig))


def sigma_buffer(timesteps: int, sigma_begin: float = 0.01, sigma_end: float = 1.0) -> torch.Tensor:
    s = torch.FloatTensor(np.exp(np.linspace(np.log(sigma_begin), np.log(sigma_end), timesteps)))
    return torch.cat([torch.zeros(1), s], dim=0)


def sigma_norm_monte_carlo(sigmas_1_to_T: torch.Tensor, sn: int = 10000, seed: Optional[int] = None) -> torch.Tensor:
    """Fresh estimate of E[score^2] of the wrapped normal, used only when no
    checkpoint buffer is available (random-init benchmarking)."""
    g = torch.Generator().manual_seed(0 if seed is None else seed)
    s = sigmas_1_to_T[None, :].repeat(sn, 1)
    x = (sigmas_1_to_T * torch.randn(s.shape, generator=g)) % 1.0
    num = 0
    den = 0
    for i in range(-10, 11):
        gq = torch.exp(-((x + i) ** 2) / 2 / s ** 2)
        num = num + (x + i) / s ** 2 * gq
        den = den + gq
    return torch.cat([torch.ones(1), ((num / den) ** 2).mean(dim=0)])


def d3pm_closed_form(betas: torch.Tensor):
    """The three distinct entries of Q_s and of Qbar_s for s = 0..T (fp32, same
    multiply order as the reference's matmul chain, diff_utils.py:176-183)."""
    T1 = betas.shape[0]
    one = torch.ones((), dtype=torch.float32)
    omb = (one - betas).to(torch.float32)           # Q_s[c,c]
    q00 = (omb + betas).to(torch.float32)           # Q_s[0,0]
    diag = torch.empty(T1, dtype=torch.float32)
    off = torch.empty(T1, dtype=torch.float32)
    qb00 = torch.empty(T1, dtype=torch.float32)
    diag[0], off[0], qb00[0] = omb[0], betas[0], q00[0]
    for s in range(1, T1):
        # (Qbar_{s-1} Q_s)[c,c] = diag*omb ; [c,0] = off*q00 + diag*beta ; [0,0] = qb00*q00
        off[s] = off[s - 1] * q00[s] + diag[s - 1] * betas[s]
        diag[s] = diag[s - 1] * omb[s]
        qb00[s] = qb00[s - 1] * q00[s]
    return omb, q00, diag, off, qb00


def coefficient_table(timesteps: int, sigmas_norm: torch.Tensor, step_lr: float,
                      sigma_begin: float = 0.01, sigma_end: float = 1.0, mode: str = "cosine",
                      q_mats: Optional[torch.Tensor] = None,
                      q_one_step_mats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[T+1, 16] fp32.  Row t holds every scalar the update at timestep t needs.

    If the checkpoint's `d3pm.q_mats` / `d3pm.q_one_step_mats` buffers are given,
    the D3PM entries are read from them (bit-identical to what the reference
    gathers); otherwise they come from the closed-form recurrence."""
    bb = beta_buffers(timesteps, mode)
    sx = sigma_buffer(timesteps, sigma_begin, sigma_end)
    sn = sigmas_norm.to(torch.float32).cpu()
    tab = torch.zeros(timesteps + 1, C_NCOLS, dtype=torch.float32)
    omb, q00, diag, off, qb00 = d3pm_closed_form(bb["betas"])
    if q_one_step_mats is not None:
        q1 = q_one_step_mats.to(torch.float32).cpu()
        omb, q00 = q1[:, 1, 1].clone(), q1[:, 0, 0].clone()
        beta_q = q1[:, 1, 0].clone()
    else:
        beta_q = bb["betas"]
    if q_mats is not None:
        qm = q_mats.to(torch.float32).cpu()
        diag, off, qb00 = qm[:, 1, 1].clone(), qm[:, 1, 0].clone(), qm[:, 0, 0].clone()
    for t in range(1, timesteps + 1):
        a = bb["alphas"][t]
        ac = bb["alphas_cumprod"][t]
        tab[t, C_LAT_C0] = 1.0 / torch.sqrt(a)
        tab[t, C_LAT_C1] = (1 - a) / torch.sqrt(1 - ac)
        tab[t, C_LAT_SIG] = bb["sigmas"][t]
        s_t, s_p = sx[t], sx[t - 1]
        tab[t, C_X_STEP] = s_t ** 2 - s_p ** 2
        tab[t, C_X_STD] = torch.sqrt((s_p ** 2 * (s_t ** 2 - s_p ** 2)) / (s_t ** 2))
        tab[t, C_X_SQRT_NORM] = torch.sqrt(sn[t])
        st2 = step_lr * (s_t / sigma_begin) ** 2
        tab[t, C_X_STEP2] = st2
        tab[t, C_X_STD2] = torch.sqrt(2 * st2)
        tab[t, C_D_BETA] = beta_q[t - 1]
        tab[t, C_D_OMB] = omb[t - 1]
        tab[t, C_D_Q00] = q00[t - 1]
        k = t - 2  # python negative index at t=1 wraps exactly like the reference (masked anyway)
        tab[t, C_D_DIAG] = diag[k]
        tab[t, C_D_OFF] = off[k]
        tab[t, C_D_QB00] = qb00[k]
    return tab


def time_embedding_table(timesteps: int, dim: int) -> torch.Tensor:
    """SinusoidalTimeEmbeddings for t = 0..T -> [T+1, dim] (cspnet.py:28-35)."""
    half = dim // 2
    scale = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half) * -scale)
    t = torch.arange(timesteps + 1)
    ang = t[:, None] * freqs[None, :]
    return torch.cat((ang.sin(), ang.cos()), dim=-1)
