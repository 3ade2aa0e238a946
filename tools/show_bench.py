#!/usr/bin/env python3
"""Print the key numbers of a bench.py JSON line.   python tools/show_bench.py gpurun_out/x.json"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.strip().startswith("{")][-1])      # (NCCL may print its version line first)
r = d.get("roofline") or {}
print("value %.0f %s | dominant %s frac %.3f | pair frac %.3f | e2e %s | clocks %s" % (
    d["value"], d["unit"], r.get("kernel"), r.get("frac", 0), d.get("frac_of_hbm_peak_pair") or 0,
    (d.get("e2e") or {}).get("value"), d.get("clocks")))
print("copy_probe", (d.get("e2e") or {}).get("copy_probe"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
for k, v in (d.get("extras") or {}).items():
    if isinstance(v, dict):
        print("  %-48s %8.3f ms  %5.1f %% of HBM peak  %s" % (k, v.get("ms_per_pair", 0), 100 * v.get("frac_of_hbm_peak", 0),
              {a: b for a, b in v.items() if a in ("columns", "images", "images_per_gpu", "signals", "fp32_tflops")}))
    else:
        print(" ", k, v)
