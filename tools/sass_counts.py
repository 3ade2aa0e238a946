#!/usr/bin/env python3
"""Native-code evidence: per kernel family, the SASS mnemonics that prove the Blackwell data path (B200_PROFILING.md,
"What proves a Blackwell-native kernel") counted from `cuobjdump -sass` of the built library, next to the `-Xptxas -v`
resource lines of the same build.     python tools/sass_counts.py > profiles/r02_sass_counts.md
    UBLKCP   = cp.async.bulk (1-D TMA bulk copy)          UTMALDG = cp.async.bulk.tensor (tensor-map TMA load)
    UBLKPF   = cp.async.bulk.prefetch.L2 (software prefetch of a future tile into L2)
    LDGSTS   = cp.async (16-byte global -> shared, the one-pass 3-D loader: periodic wrap per chunk)
    SYNCS    = mbarrier arrive / try_wait                  LDS.128 / STS.128 = 16-byte shared-memory accesses
    STG.E.128 (+ .EF = evict-first) = 16-byte global stores;   LDG.E.128 = 16-byte global loads;   FFMA / DFMA = the arithmetic"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "wavelets.jl_b200", "lib", "libwavelets_b200.so")
PATS = [("UBLKCP", r"\bUBLKCP"), ("UBLKPF", r"\bUBLKPF"), ("UTMALDG", r"\bUTMALDG"), ("SYNCS", r"\bSYNCS"), ("LDGSTS", r"\bLDGSTS"), ("LDS.128", r"\bLDS\.128"), ("STS.128", r"\bSTS\.128"),
        ("LDG.E.128", r"\bLDG\.E\.128"), ("STG.E.128", r"\bSTG\.E\.128"), ("STG.E.EF.128", r"\bSTG\.E\.EF\.128"),
        ("STG.E.EF.64", r"\bSTG\.E\.EF\.64"), ("FFMA", r"\bFFMA"), ("DFMA", r"\bDFMA"), ("BAR.SYNC", r"\bBAR\.SYNC")]


def family(name):
    m = re.match(r"(?:void )?(?:[A-Za-z0-9_]+::|\(anonymous namespace\)::)*(k_[A-Za-z0-9_]+)", name)
    return m.group(1) if m else name[:40]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    dem = {}
    fam = collections.OrderedDict()
    cur = None
    names = re.findall(r"Function : (\S+)", out)
    filt = subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    dem = dict(zip(names, filt))
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = family(dem.get(m.group(1), m.group(1)))
            fam.setdefault(cur, {"kernels": 0, **{k: 0 for k, _ in PATS}})["kernels"] += 1
            continue
        if cur is None or "/*" not in line:
            continue
        for k, pat in PATS:
            if re.search(pat, line):
                fam[cur][k] += 1
    print("# SASS evidence (round 2): `cuobjdump -sass wavelets.jl_b200/lib/libwavelets_b200.so`, counted per kernel family\n")
    print("Counts are summed over every template instantiation of the family (`kernels` column: element types, filter lengths,")
    print("strict / fast).  Mnemonic key: tools/sass_counts.py.\n")
    hdr = ["kernel family", "kernels"] + [k for k, _ in PATS]
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
    tot = {k: 0 for k, _ in PATS}
    for f, c in fam.items():
        print("| `" + f + "` | " + str(c["kernels"]) + " | " + " | ".join(str(c[k]) for k, _ in PATS) + " |")
        for k, _ in PATS:
            tot[k] += c[k]
    print("| **total** | " + str(sum(c["kernels"] for c in fam.values())) + " | " + " | ".join(str(tot[k]) for k, _ in PATS) + " |")
    # ptxas -v: registers / shared memory / spills of the headline instantiations
    print("\n## `-Xptxas -v` resource lines (wavelets.jl_b200/lib/obj/*.ptxas.log), headline instantiations\n")
    want = [r"k_ana_tiles<float, 8, false>", r"k_syn_tiles<float, 8, false>", r"k_ana_tiles<double, 8, false>", r"k_syn_tiles<double, 8, false>",
            r"k_lift1d_(ana|syn)<float, wb::ShapeCdf97\w, false", r"k_lift2d_fwd_tma<float, wb::ShapeCdf97F, false", r"k_lift2d_inv_tma<float, wb::ShapeCdf97I, false",
            r"k_lift2d_fwd_tma<float, wb::ShapeFirA<8>, false", r"k_fir3d_\w+<float, 12, false", r"k_lift2d_tailfast<float, wb::ShapeCdf97F, false", r"k_wpt_sub_ana<float, 16, false", r"k_wpt_sub_syn<float, 16, false",
            r"k_pkt_ana<float, 16, false", r"k_pkt_syn<float, 16, false"]
    print("| kernel | registers | spill stores / loads | static smem |")
    print("|---|---|---|---|")
    for log in sorted(glob.glob(os.path.join(ROOT, "wavelets.jl_b200", "lib", "obj", "*.ptxas.log"))):
        txt = open(log).read()
        blocks = re.split(r"ptxas info\s+: Compiling entry function '", txt)
        for b in blocks[1:]:
            mang = b.split("'")[0]
            d = subprocess.run(["cu++filt", mang], capture_output=True, text=True).stdout.strip()
            d = d.replace("(int)", "").replace("(bool)0", "false").replace("(bool)1", "true")
            if not any(re.search(w, d) for w in want):
                continue
            regs = re.search(r"Used (\d+) registers", b)
            sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
            sm = re.search(r"(\d+) bytes smem", b)
            print(f"| `{d[:150]}` | {regs.group(1) if regs else '?'} | {sp.group(1) + ' / ' + sp.group(2) if sp else '?'} | {sm.group(1) if sm else '0'} |")


if __name__ == "__main__":
    main()
