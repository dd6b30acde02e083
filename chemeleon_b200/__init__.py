"""chemeleon_b200 -- B200-native sampler for Chemeleon's text-conditioned crystal diffusion."""
from .config import SamplerConfig  # noqa: F401

__all__ = ["SamplerConfig"]
