#!/usr/bin/env python3
"""Regression pins for the parts of the oracle the reference holds no vectors for (MODWT, denoise, best basis): outputs of THIS
repository's oracle on fixed inputs, so that an accidental change of the restatement shows up as a diff.  They are NOT upstream
golden vectors (those are wavelab_golden.json) and say nothing about parity with the reference beyond what the oracle's own
pinning says (tests/test_oracle.py)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle as orc
import wavelets_b200 as wb


def signal(n):
    t = np.linspace(0, 1, n)
    return np.sqrt(t * (1 - t)) * np.sin(2 * np.pi * 1.05 / (t + 0.05)) + 0.05 * np.cos(37.0 * np.arange(n) ** 1.1)


def main():
    x = signal(64)
    db4, sym5 = wb.wavelet(wb.WT.db4), wb.wavelet(wb.WT.sym5)
    cdf = wb.wavelet(wb.WT.cdf97, wb.WT.Lifting)
    out = {"x64": x.tolist(),
           "modwt_db4_L3": orc.modwt(x, np.asarray(db4.qmf), 3).ravel(order="F").tolist(),
           "noisest_sym5": orc.noisest(x, sym5),
           "denoise_sym5_L4": orc.denoise(x, sym5, 4).tolist(),
           "denoise_sym5_L4_TI": orc.denoise(x, sym5, 4, TI=True).tolist(),
           "denoise_cdf97_soft_TI": orc.denoise(x, cdf, 4, kind="soft", TI=True, nspin=4).tolist()}
    best, bf, af = orc.bestbasistree(x, db4, wb.maketree(64, None, "full"))
    out["bestbasis_db4_tree"] = best.tolist()
    out["bestbasis_db4_entr_bf"] = bf.tolist()
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "derived_pins.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
