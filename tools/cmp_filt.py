import os, sys, torch
sys.path.insert(0, "/root/repo")
import wavelets_b200 as wb
B = 4096
wl = wb.wavelet(wb.WT.db4)
x = torch.randn((B, 1 << 20), device="cuda").t()
y = wb.dwtc(x, wl)
b = 2 * 4 * B * (1 << 20) / 1e9
def timeit(fn, reps=6):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def run(tag, env):
    for k in ("WB200_TILE_F32", "WB200_TILE_F32_INV", "WB200_F1D_NT", "WB200_F1D_NT_INV", "WB200_KMAX", "WB200_TAILMAX_F32", "WB200_TAILMAX_F32_INV"):
        os.environ.pop(k, None)
    os.environ.update(env)
    f = timeit(lambda: wb.dwtc(x, wl)); i = timeit(lambda: wb.idwtc(y, wl))
    print(f"{tag:50s} fwd {f:7.3f} ms {b/f*1e3:6.0f} GB/s  inv {i:7.3f} ms {b/i*1e3:6.0f} GB/s  pair {b*2/(f+i)*1e3:6.0f} GB/s", flush=True)
run("defaults", {})
for tf, ntf, ti, nti, k in ((2048, 64, 2048, 96, 4), (2048, 64, 2048, 96, 3), (4096, 64, 4096, 96, 4), (4096, 96, 4096, 96, 5), (2048, 64, 4096, 96, 4), (4096, 64, 2048, 96, 3)):
    for tm in (32768, 4096, 2048):
        run(f"tile {tf}/{ti} nt {ntf}/{nti} kmax {k} tailmax {tm}", {"WB200_TILE_F32": str(tf), "WB200_TILE_F32_INV": str(ti), "WB200_F1D_NT": str(ntf), "WB200_F1D_NT_INV": str(nti), "WB200_KMAX": str(k), "WB200_TAILMAX_F32": str(tm), "WB200_TAILMAX_F32_INV": str(tm)})
run("defaults again", {})
