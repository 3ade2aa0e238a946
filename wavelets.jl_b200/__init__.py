"""wavelets_b200 -- B200-native forward/inverse DWT hot path behind the Wavelets.jl API surface."""
from . import wt as WT
from .wt import wavelet

__all__ = ["WT", "wavelet"]
