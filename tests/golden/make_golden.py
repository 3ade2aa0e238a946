#!/usr/bin/env python3
"""Consolidate the reference's known-answer vectors into one committed fixture.

Source of truth: /root/reference/test/data/*.txt (WaveLab850 FWT_PO/FWT2_PO outputs
written by test/data/make_filter_data.m, plus the PyWavelets non-square Haar case
written by test/data/make_filter_data.py).  Those are the vectors the reference's
own test-suite checks `dwt` against (test/transforms.jl:2-55).

/root/reference does not exist on the GPU box, so the numbers are re-serialised
here into tests/golden/wavelab_golden.json (decimal strings are parsed to float64
and written back with repr(), which round-trips exactly).

Run once in the build container:  python tests/golden/make_golden.py
"""
import json
import os
import re
import sys

import numpy as np

REF = "/root/reference/test/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wavelab_golden.json")

# file stem -> (wavelet class name in the reference, vanishing moments / number)
# mapping documented at test/transforms.jl:3-7
NAME_MAP = {}
for taps in range(4, 22, 2):
    NAME_MAP[f"Daubechies{taps}"] = f"db{taps // 2}"
# Coiflet1.txt exists on disk but the reference test never reads it (wnum Coiflet = 2:5) and it is NOT the
# reference's coif2 table (different published coefficients), so it is not a golden vector for this path.
for num, vm in zip((2, 3, 4, 5), (4, 6, 8, 10)):
    NAME_MAP[f"Coiflet{num}"] = f"coif{vm}"
NAME_MAP["Haar0"] = "haar"
for vm in range(4, 11):
    NAME_MAP[f"Symmlet{vm}"] = f"sym{vm}"
for num, vm in zip((1, 3, 5), (2, 4, 6)):
    NAME_MAP[f"Battle{num}"] = f"batt{vm}"
NAME_MAP["Vaidyanathan0"] = "vaid"
NAME_MAP["Beylkin0"] = "beyl"


def load(path):
    return np.loadtxt(path, dtype=np.float64)


def main():
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not present: this script only runs in the build container")
    gold = {
        "_source": "JuliaDSP/Wavelets.jl test/data (WaveLab850 + PyWavelets), see make_golden.py",
        "data1d": load(os.path.join(REF, "filter1d_data.txt")).tolist(),
        "data2d": load(os.path.join(REF, "filter2d_data.txt")).tolist(),
        "nonsquare_data": load(os.path.join(REF, "filter2d_nonsquare_data.txt")).tolist(),
        "nonsquare_haar_L1": load(os.path.join(REF, "filter2d_nonsquare_Haar0.txt")).tolist(),
        "expected1d": {},
        "expected2d": {},
    }
    for fn in sorted(os.listdir(REF)):
        m = re.match(r"filter(1d|2d)_([A-Za-z]+\d+)\.txt$", fn)
        if not m or m.group(2) not in NAME_MAP:
            continue
        key = NAME_MAP[m.group(2)]
        gold["expected" + m.group(1)][key] = load(os.path.join(REF, fn)).tolist()
    assert len(gold["expected1d"]) == len(gold["expected2d"]) == 26, (
        len(gold["expected1d"]), len(gold["expected2d"]))
    with open(OUT, "w") as f:
        json.dump(gold, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(gold["expected1d"]), "wavelets")


if __name__ == "__main__":
    main()
