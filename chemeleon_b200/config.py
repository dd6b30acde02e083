"""Architecture/sampler hyper-parameters the kernels are compiled for.

Mirrors the keys the reference sampler reads from its sacred `_config`
(chemeleon/modules/chemeleon.py:36-91; defaults chemeleon/config.py:28-42,52-55).
The CUDA kernels are specialised for the default shape family and refuse
anything else -- there is no generic fallback.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Any, Dict


@dataclass(frozen=True)
class SamplerConfig:
    hidden_dim: int = 512
    time_dim: int = 128
    text_dim: int = 512
    max_atoms: int = 104
    num_layers: int = 6
    num_freqs: int = 128
    timesteps: int = 1000
    act_fn: str = "silu"
    dis_emb: str = "sin"
    edge_style: str = "fc"
    ln: bool = True
    ip: bool = True
    smooth: bool = False
    pred_atom_types: bool = True
    text_guide: bool = True
    beta_schedule: str = "cosine"
    sigma_begin: float = 0.01
    sigma_end: float = 1.0

    @classmethod
    def from_hparams(cls, hp: Any) -> "SamplerConfig":
        """Build from `model.hparams` / the checkpoint's `hyper_parameters` dict."""
        if not isinstance(hp, dict):
            hp = dict(vars(hp)) if hasattr(hp, "__dict__") else dict(hp)
        names = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in hp.items() if k in names})

    def validate(self) -> None:
        """The sm_100a kernels exist for exactly this shape family."""
        want: Dict[str, Any] = dict(hidden_dim=512, time_dim=128, text_dim=512, max_atoms=104,
                                    num_freqs=128, act_fn="silu", dis_emb="sin", edge_style="fc",
                                    ln=True, ip=True, smooth=False, pred_atom_types=True,
                                    beta_schedule="cosine")
        bad = {k: getattr(self, k) for k, v in want.items() if getattr(self, k) != v}
        if bad:
            raise ValueError(
                f"chemeleon_b200 kernels are specialised for {want}; unsupported: {bad} "
                "(edge_style='knn' is broken in the reference itself, SURVEY.md App. A)")
        if not (1 <= self.num_layers <= 16):
            raise ValueError("num_layers out of range")

    @property
    def dis_dim(self) -> int:
        return 2 * 3 * self.num_freqs

    @property
    def edge_in_dim(self) -> int:
        return 2 * self.hidden_dim + 9 + self.dis_dim
