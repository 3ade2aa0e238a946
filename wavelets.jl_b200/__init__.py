"""wavelets_b200 -- B200-native (sm_100a) forward/inverse DWT hot path behind the Wavelets.jl API surface.

    import wavelets_b200 as wb                    # loads ./wavelets.jl_b200/
    wt = wb.wavelet(wb.WT.db4)                    # wavelet(WT.db4)
    y  = wb.dwt(x_cuda, wt)                       # dwt(x, wt)          (x: torch CUDA tensor, column-major)
    x2 = wb.idwt(y, wt)
    yb = wb.dwtc(xb_cuda, wt)                     # column-wise batch: last dim = independent signals

Layers: `WT` (wavelet descriptors, host only) -> `transforms` (API mirror of src/Transforms/transforms_main.jl)
-> C ABI `lib/libwavelets_b200.so` (include/wavelets_b200.h) -> hand-written CUDA kernels (csrc/).
"""
from . import wt as WT
from . import util as Util
from .wt import wavelet
from .util import maxtransformlevels, maketree, isvalidtree, detailindex, detailrange, detailn
from .transforms import (modwt, imodwt, maxmodwttransformlevels, dwt, idwt, dwt_, idwt_, wpt, iwpt, wpt_, iwpt_, dwtc, idwtc, dwt_oop_, idwt_oop_,
                         ArgumentError, DimensionMismatch, set_strict_fp, colmajor, release_scratch)
from . import threshold as Threshold
from .threshold import (HardTH, SoftTH, SemiSoftTH, SteinTH, BiggestTH, PosTH, NegTH, threshold, threshold_, VisuShrink, denoise, noisest,
                        ShannonEntropy, LogEnergyEntropy, coefentropy, bestbasistree)

__all__ = ["modwt", "imodwt", "maxmodwttransformlevels", "WT", "Util", "wavelet", "maxtransformlevels", "maketree", "isvalidtree", "detailindex",
           "detailrange", "detailn", "dwt", "idwt", "dwt_", "idwt_", "wpt", "iwpt", "wpt_", "iwpt_",
           "dwtc", "idwtc", "dwt_oop_", "idwt_oop_", "ArgumentError", "DimensionMismatch",
           "set_strict_fp", "colmajor", "release_scratch", "Threshold", "HardTH", "SoftTH", "SemiSoftTH", "SteinTH", "BiggestTH", "PosTH", "NegTH",
           "threshold", "threshold_", "VisuShrink", "denoise", "noisest", "ShannonEntropy", "LogEnergyEntropy", "coefentropy", "bestbasistree"]
