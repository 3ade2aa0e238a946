#!/usr/bin/env python3
"""Sweep the fused 1-D kernels' tuning knobs (environment variables read by csrc/fused1d.cu at dispatch time) and
print per-kernel device times from the library's profiling hook.  Run on the GPU box:
    python tools/tune_fused1d.py --dtype f32 --batch 2048
"""
import argparse, ctypes as C, itertools, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import wavelets_b200 as wb
from wavelets_b200 import _lib


def run(L, x, y, ws, wsb, dims, B, qp, flen, lv, code, reps=3):
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    def step():
        assert L.wb200_dwt_filter(y.data_ptr(), x.data_ptr(), 1, dims, B, qp, flen, lv, 1, code, ws.data_ptr(), wsb, sp, 0) == 0, L.wb200_last_error_string()
        assert L.wb200_dwt_filter(x.data_ptr(), y.data_ptr(), 1, dims, B, qp, flen, lv, 0, code, ws.data_ptr(), wsb, sp, 0) == 0, L.wb200_last_error_string()
    step(); torch.cuda.synchronize()
    L.wb200_profile_enable(1)
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    L.wb200_profile_enable(0)
    buf = C.create_string_buffer(1 << 14)
    nb = L.wb200_profile_collect(buf, len(buf))
    out = {}
    for ln in buf.raw[:nb].decode().splitlines():
        nm, cnt, ms = ln.split()
        out[nm] = float(ms) / reps          # per direction-pass (a pass may be several launches)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--wavelet", default="db4")
    a = ap.parse_args()
    L = _lib.lib()
    dev = torch.device("cuda:0")
    tdt = torch.float32 if a.dtype == "f32" else torch.float64
    esz = 4 if a.dtype == "f32" else 8
    code = _lib.F32 if a.dtype == "f32" else _lib.F64
    wt = wb.wavelet(getattr(wb.WT, a.wavelet))
    q = np.ascontiguousarray(wt.qmf); qp = q.ctypes.data_as(C.POINTER(C.c_double))
    n, B = a.n, a.batch
    lv = wb.maxtransformlevels(n)
    x = torch.randn((B, n), dtype=tdt, device=dev); y = torch.empty_like(x)
    dims = _lib.dims_array([n])
    ws = torch.empty(max(4 * n * B * esz // 16, 1 << 20), dtype=torch.uint8, device=dev)
    sfx = "F32" if a.dtype == "f32" else "F64"
    tiles = [4096, 8192, 16384] if a.dtype == "f32" else [2048, 4096, 8192]
    tails = [8192, 16384, 32768] if a.dtype == "f32" else [4096, 8192, 16384]
    bytes_pass = 2.0 * esz * n * B
    for tile, tail, hdiv in itertools.product(tiles, tails, [4, 16]):
        os.environ[f"WB200_TILE_{sfx}"] = str(tile)
        os.environ[f"WB200_TAILMAX_{sfx}"] = str(tail)
        os.environ["WB200_HALO_DIV"] = str(hdiv)
        try:
            t = run(L, x, y, ws, ws.numel(), dims, B, qp, len(q), lv, code)
        except AssertionError as e:
            print(json.dumps({"tile": tile, "tail": tail, "halo_div": hdiv, "error": str(e)[:100]})); continue
        fwd = sum(v for k, v in t.items() if "ana" in k); inv = sum(v for k, v in t.items() if "syn" in k)
        print(json.dumps({"tile": tile, "tail": tail, "halo_div": hdiv, "ms": {k: round(v, 3) for k, v in t.items()},
                          "fwd_GBs": round(bytes_pass / fwd / 1e6, 1), "inv_GBs": round(bytes_pass / inv / 1e6, 1),
                          "pair_GBs": round(2 * bytes_pass / (fwd + inv) / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
