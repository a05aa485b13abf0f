"""pd_mg_pin_corrosion_b200 -- B200-native peridynamic bond-summation hot path.

Host-side mirror (Python, ctypes) of the reference's Config / Grid / Fields /
solver / coupling-loop surface over the C-ABI CUDA library `libpdgpu.so`
(include/pdgpu.h).  There is no CPU fallback: every operator raises if the CUDA
library is missing.
"""
from .config import Config, PdConfig  # noqa: F401

__all__ = ["Config", "PdConfig"]
