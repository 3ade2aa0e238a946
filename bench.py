#!/usr/bin/env python3
"""bench.py -- Msamples/s of the dwt+idwt pair on B200, next to the CPU reference path.

Contract (see task brief):  python bench.py --gpus N --steps K --warmup W [--impl reference]
  * N = 1: BASELINE.json configs[1]: 1-D filter-bank dwt + idwt, WT.db4, Float32, N = 2^20, L = 20, a batch of
    independent columns resident in HBM (batch >> L2, so no L2 flush is needed between iterations).
  * N > 1 (torchrun, one rank per GPU): every rank owns an equal block of columns (weak scaling, no data-path
    collective); value = all columns of all ranks / max-over-ranks device time.
  * one "step" = dwt of the whole batch followed by idwt of the result (the metric is quoted on the pair).
  * value          : device-resident throughput, CUDA-event timed on the launching stream.
  * e2e            : same metric through the C ABI's host-buffer entry points (pinned host memory, H2D + D2H
                     inside the timed region).
  * roofline       : the dominant kernel's algorithmic bytes (2*sizeof(T) per sample per direction)
                     / its CUDA-event time inside the timed region, against MEASURED_PEAKS.json hbm_gbs.
  * cpu_baseline   : the CPU oracle (a restatement of the reference's algorithm; "port") on a bounded sample,
                     all host cores (OpenMP over columns) -- rank 0, N = 1 only.
  * --impl reference: times that CPU path alone and prints the same JSON line with "impl": "reference".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_SIGNAL = 1 << 20
LEVELS = 20
METRIC = "Msamples/s dwt+idwt (1-D db4 N=2^20, 2-D cdf97 4096\u00b2); % HBM roofline"   # BASELINE.json "metric", verbatim


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="columns per GPU (0 = auto from free HBM, <= 8192)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--e2e-batch", type=int, default=512)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs (2-D, 3-D, packets, the other dtype)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks line")
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a thread (10 ms period; started before the
    warm-up so that it is already running when the timed steps begin, samples outside [mark_begin, mark_end] are dropped),
    `nvidia-smi -lms` as the fallback when NVML is not importable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
            0x80: "hw_power_brake_slowdown"}

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.nv = index, [], None, None
        self.samples, self.stop_flag, self.t0, self.t1, self.smax = [], False, None, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nv = (pynvml, h)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        pynvml, h = self.nv
        while not self.stop_flag:
            try:
                mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    bits = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    bits = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.samples.append((time.perf_counter(), mhz, bits))
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def _inside(self, t):
        return (self.t0 is None or t >= self.t0) and (self.t1 is None or t <= self.t1)

    def stop(self):
        self.stop_flag = True
        sm, smax, reasons = [], self.smax, set()
        if self.nv is not None:
            for t, mhz, bits in self.samples:
                if not self._inside(t):
                    continue
                sm.append(mhz)
                for bit, nm in self.BITS.items():
                    if bits & bit:
                        reasons.add(nm)
            src = "nvml"
        elif self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for t, r in self.rows:
                f = [x.strip() for x in r.split(",")]
                if len(f) < 7 or not self._inside(t):
                    continue
                try:
                    sm.append(float(f[0])); smax = float(f[1])
                except ValueError:
                    continue
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            src = "nvidia-smi"
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (NVML and nvidia-smi unavailable)"], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "source": src}


# ------------------------------------------------------------------------------------------------------
# CPU reference path (the oracle port of the reference's algorithm), all host cores
# ------------------------------------------------------------------------------------------------------
def cpu_pair_rate(dtype_np, seconds, qmf, n=N_SIGNAL, L=LEVELS):
    """Time dwt+idwt of as many columns as fit in ~`seconds` on all host cores; returns (Msamples/s, cols, cores)."""
    import numpy as np
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(42)
    probe = np.asfortranarray(rng.standard_normal((n, cores)).astype(dtype_np))
    t0 = time.perf_counter()
    y = orc.dwt_filter_batch(probe, 1, qmf, L, True, cores)
    orc.dwt_filter_batch(y, 1, qmf, L, False, cores)
    t_probe = time.perf_counter() - t0
    reps = max(1, int(seconds / max(t_probe, 1e-3)))
    cols = cores * min(reps, 16)
    x = np.asfortranarray(rng.standard_normal((n, cols)).astype(dtype_np))
    t0 = time.perf_counter()
    y = orc.dwt_filter_batch(x, 1, qmf, L, True, cores)
    orc.dwt_filter_batch(y, 1, qmf, L, False, cores)
    dt = time.perf_counter() - t0
    return n * cols / dt / 1e6, cols, cores


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, restated (oracle port), on the host cores."""
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import wavelets_b200 as wb
    qmf = wb.wavelet(wb.WT.db4).qmf
    dt_np = np.float32 if args.dtype == "f32" else np.float64
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    cols = cores * 2
    rng = np.random.default_rng(42)
    x = np.asfortranarray(rng.standard_normal((N_SIGNAL, cols)).astype(dt_np))

    def step():
        y = orc.dwt_filter_batch(x, 1, qmf, LEVELS, True, cores)
        orc.dwt_filter_batch(y, 1, qmf, LEVELS, False, cores)
    for _ in range(min(args.warmup, 2)):
        step()
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    val = N_SIGNAL * cols / dt / 1e6
    sample = f"{cols} columns of N=2^20 per step (dwt+idwt), OpenMP over columns"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": "1-D db4 filter-bank dwt+idwt, N=2^20, L=20, CPU oracle port of src/Transforms "
                               "(Julia reference not runnable here: no julia binary)", "columns": cols},
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def load_peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    return peak, src


def collect_kernels(L):
    buf = C.create_string_buffer(1 << 14)
    nb = L.wb200_profile_collect(buf, len(buf))
    kern = {}
    for line in buf.raw[:nb].decode().splitlines():
        nm, cnt, ms = line.split()
        kern[nm] = (int(cnt), float(ms))
    return kern


def columns_that_fit(L, _lib, dev, esz, code, cap=8192):
    """As many columns of N=2^20 as comfortably fit (x, y and the library workspace), at most BASELINE's 8192."""
    import torch
    free_b, _ = torch.cuda.mem_get_info(dev)
    B = cap
    while B > 64:
        need = 2 * N_SIGNAL * B * esz + L.wb200_workspace_bytes(0, 1, _lib.dims_array([N_SIGNAL]), B, LEVELS, code, 0)
        if need < 0.80 * free_b:
            break
        B //= 2
    return B


def measure_1d(L, _lib, dev, stream, qmf, dtype_name, B, steps, warmup, barrier, seed, sampler=None):
    """configs[1]: dwt + idwt of B resident columns of N = 2^20 (db4, L = 20), CUDA events on the launching stream,
    per-kernel device times from the library's own event hook inside the timed region."""
    import torch
    tdt = torch.float32 if dtype_name == "f32" else torch.float64
    code = _lib.F32 if dtype_name == "f32" else _lib.F64
    esz = 4 if dtype_name == "f32" else 8
    qp = qmf.ctypes.data_as(C.POINTER(C.c_double))
    dims = _lib.dims_array([N_SIGNAL])
    ws_bytes = L.wb200_workspace_bytes(0, 1, dims, B, LEVELS, code, 0)
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(seed)
    x = torch.empty((B, N_SIGNAL), dtype=tdt, device=dev)
    for b0 in range(0, B, 256):                                   # randn in slabs: no 2x temporary
        x[b0:b0 + 256].normal_(generator=gen)
    y = torch.empty_like(x)
    sp = C.c_void_p(stream.cuda_stream)

    def step():
        rc = L.wb200_dwt_filter(y.data_ptr(), x.data_ptr(), 1, dims, B, qp, len(qmf), LEVELS, 1, code,
                                ws.data_ptr(), ws_bytes, sp, 0)
        assert rc == 0, L.wb200_last_error_string()
        rc = L.wb200_dwt_filter(x.data_ptr(), y.data_ptr(), 1, dims, B, qp, len(qmf), LEVELS, 0, code,
                                ws.data_ptr(), ws_bytes, sp, 0)
        assert rc == 0, L.wb200_last_error_string()

    for _ in range(warmup):
        step()
    barrier()
    L.wb200_launch_count(1)
    L.wb200_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.mark_begin()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    barrier()
    if sampler:
        sampler.mark_end()
    L.wb200_profile_enable(0)
    ms_total = e0.elapsed_time(e1)
    launches = int(L.wb200_launch_count(1))
    kern = collect_kernels(L)
    res = {"ms_step": ms_total / steps, "ms_total": ms_total, "launches": launches, "kern": kern, "B": B, "esz": esz, "x": x, "y": y}
    return res


def roofline_block(kern, ms_total, steps, esz, B, dtype_name, peak, peak_src):
    """Dominant kernel of the timed region: algorithmic bytes (2*sizeof(T) per sample per direction pass) over its
    CUDA-event time, against the measured HBM peak."""
    if not kern:
        return None
    dom = max(kern, key=lambda k: kern[k][1])
    cnt, ms = kern[dom]
    # every kernel name belongs to one direction (analysis / synthesis), so over the timed region it covered `steps`
    # direction-passes; one pass moves 2*esz bytes per sample algorithmically, whatever the number of launches it is split into
    alg_bytes_per_pass = 2.0 * esz * N_SIGNAL * B
    ach = alg_bytes_per_pass * steps / (ms * 1e-3) / 1e9
    traffic, tsrc = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        per_col = tj.get(dtype_name, {}).get(dom, {}).get("dram_bytes_per_column")
        if per_col:
            traffic = per_col * B
            tsrc = ("STATIC: dram__bytes_read.sum + dram__bytes_write.sum per column from the committed ncu --set full capture "
                    f"({tj.get('_source', 'profiles/traffic.json')}), scaled to this batch; not re-measured in this run")
    except Exception:
        pass
    return {"bound": "hbm", "kernel": dom, "launches": cnt, "kernel_ms_total": ms, "share_of_step": ms / ms_total,
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_source": tsrc,
            "peak_source": peak_src, "algorithmic_bytes_per_direction_pass": alg_bytes_per_pass,
            "all_kernels_ms": {k: v[1] for k, v in kern.items()}}


def copy_probe(dev, stream, world, allreduce_sum, gib=1.0):
    """H2D-only and D2H-only bandwidth of THIS rank's pinned buffer while every rank copies at once (diagnoses what caps
    the end-to-end number: host DRAM / PCIe root / the pipeline).  Returns per-rank and summed GB/s."""
    import torch
    nb = int(gib * (1 << 30))
    h = torch.empty(nb, dtype=torch.uint8).pin_memory()
    d = torch.empty(nb, dtype=torch.uint8, device=dev)
    out = {}
    for name, (dst, src) in (("h2d", (d, h)), ("d2h", (h, d))):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        gbs = 3 * nb / (e0.elapsed_time(e1) * 1e-3) / 1e9
        out[name + "_gbs_this_rank"] = gbs
        out[name + "_gbs_all_ranks"] = allreduce_sum(gbs, dev) if world > 1 else gbs
    return out


def device_copy_probe(x, y, dev, reps=6):
    """What a plain device copy (torch's `y.copy_(x)`, the kernel MEASURED_PEAKS.json's hbm_gbs was measured with) reaches
    on the SAME operands right after the timed steps -- same footprint (read 32 GiB, write 32 GiB at 8192 Float32 columns),
    same board state (power cap, clocks).  Context for `roofline.frac`, whose denominator is the driver's burst figure."""
    import torch
    y.copy_(x); y.copy_(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(reps):
        y.copy_(x)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    nbytes = 2 * x.numel() * x.element_size()
    return {"gbs": nbytes / (ms * 1e-3) / 1e9, "ms": ms, "bytes": nbytes,
            "note": "torch y.copy_(x) on the step's own operands, back to back after the timed region (sustained, same power state)"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    import wavelets_b200 as wb
    from wavelets_b200 import _lib
    from wavelets_b200.shard import allreduce_max, allreduce_sum

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product has no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    wt = wb.wavelet(wb.WT.db4)
    qmf = np.ascontiguousarray(wt.qmf)
    qp = qmf.ctypes.data_as(C.POINTER(C.c_double))
    tdt = torch.float32 if args.dtype == "f32" else torch.float64
    code = _lib.F32 if args.dtype == "f32" else _lib.F64
    esz = 4 if args.dtype == "f32" else 8
    stream = torch.cuda.current_stream(dev)
    peak, peak_src = load_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    B = args.batch if args.batch > 0 else columns_that_fit(L, _lib, dev, esz, code)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    warm = max(args.warmup, 3)
    m = measure_1d(L, _lib, dev, stream, qmf, args.dtype, B, args.steps, warm, barrier, 42 + rank, sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    ms_step, launches, kern = m["ms_step"], m["launches"], m["kern"]
    ms_step_max = allreduce_max(ms_step, dev) if world > 1 else ms_step
    value = N_SIGNAL * B * world / (ms_step_max * 1e-3) / 1e6
    roof = roofline_block(kern, m["ms_total"], args.steps, esz, B, args.dtype, peak, peak_src)
    step_gbs = 4.0 * esz * N_SIGNAL * B / (ms_step * 1e-3) / 1e9
    x = m.pop("x"); ycp = m.pop("y")
    if roof is not None:
        try:
            roof["device_copy_probe"] = device_copy_probe(x, ycp, dev)
            roof["frac_of_copy_probe"] = roof["achieved"] / roof["device_copy_probe"]["gbs"]
        except Exception as ex:
            roof["device_copy_probe"] = {"error": str(ex)[:200]}
    del ycp

    # ---- e2e: host buffers through the C ABI (H2D + D2H inside the timed region) ----
    dims = _lib.dims_array([N_SIGNAL])
    e2e = None
    try:
        Be = min(args.e2e_batch, B)
        xh = torch.empty((Be, N_SIGNAL), dtype=tdt).pin_memory()
        yh = torch.empty((Be, N_SIGNAL), dtype=tdt).pin_memory()
        xh.copy_(x[:Be])

        def estep():
            rc = L.wb200_dwt_filter_host(yh.data_ptr(), xh.data_ptr(), 1, dims, Be, qp, len(qmf), LEVELS, 1, code, local_rank, 0)
            assert rc == 0, L.wb200_last_error_string()
            rc = L.wb200_dwt_filter_host(xh.data_ptr(), yh.data_ptr(), 1, dims, Be, qp, len(qmf), LEVELS, 0, code, local_rank, 0)
            assert rc == 0, L.wb200_last_error_string()
        estep()
        barrier()
        ke = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(ke):
            estep()
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / ke
        dt_max = allreduce_max(dt, dev) if world > 1 else dt
        e2e = {"value": N_SIGNAL * Be * world / dt_max / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": 2 * N_SIGNAL * Be * esz, "d2h_bytes_per_step": 2 * N_SIGNAL * Be * esz,
               "columns_per_gpu": Be, "link_gbs_per_gpu_each_way": 2 * N_SIGNAL * Be * esz / dt_max / 1e9,
               "note": "wb200_dwt_filter_host: pinned host buffers, chunk pipeline (H2D, transform, D2H on rotating "
               "streams); dwt copies x in / y out, idwt copies y in / x out"}
        del xh, yh
        barrier()
        try:
            e2e["copy_probe"] = copy_probe(dev, stream, world, allreduce_sum)
        except Exception as ex:
            e2e["copy_probe"] = {"error": str(ex)[:200]}
    except Exception as ex:                                       # report, never fake
        e2e = {"value": None, "unit": "Msamples/s", "error": str(ex)[:200]}
    del x
    m.clear()
    torch.cuda.empty_cache()
    wb.release_scratch()

    # ---- the other BASELINE configs (reported next to the headline, same event protocol) ----
    extras = {}
    if not args.no_extras:
        try:
            extras = run_extras(L, wb, _lib, dev, stream, args, world, rank, barrier, allreduce_max, qmf, peak, peak_src)
        except Exception as ex:
            extras = {"error": str(ex)[:300]}

    cpu = None
    if rank == 0 and world == 1:
        try:
            v, cols, cores = cpu_pair_rate(np.float32 if args.dtype == "f32" else np.float64, args.cpu_seconds, qmf)
            cpu = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port",
                   "sample": f"{cols} columns of N=2^20 (dwt+idwt), oracle port of src/Transforms, OpenMP over columns"}
        except Exception as ex:
            cpu = {"value": None, "unit": "Msamples/s", "error": str(ex)[:200]}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_step_max, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"1-D filter-bank dwt+idwt, WT.db4 (8 taps), N=2^20, L=20, {B} columns per GPU "
                                   f"({'Float32' if esz == 4 else 'Float64'}), BASELINE.json configs[1]",
                       "columns_per_gpu": B, "parallelism": f"batch-split x{world}",
                       "l2": f"inputs {2 * N_SIGNAL * B * esz / 2**30:.1f} GiB per step >> 126 MB L2 (no flush needed)"},
            "achieved_gbs_pair": step_gbs, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks, "extras": extras,
        }
        # the metric names two legs: `value` is the 1-D db4 leg (configs[1], the one the timed steps run); the 2-D cdf97
        # 4096^2 leg (configs[2]) is measured right after it with the same event timing and reported beside it
        leg2 = extras.get("dwt2_cdf97_lifting_4096x4096_f32_L8") if isinstance(extras, dict) else None
        if leg2:
            out["value_2d_cdf97"] = {"value": leg2["msamples_per_s_pair"], "unit": "Msamples/s", "images": leg2.get("images"),
                                     "achieved_gbs_pair": leg2["achieved_gbs_pair"], "frac_of_hbm_peak": leg2["frac_of_hbm_peak"]}
        out["frac_of_hbm_peak_pair"] = step_gbs / roof["peak"] if roof and roof.get("peak") else None
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_extras(L, wb, _lib, dev, stream, args, world, rank, barrier, allreduce_max, qmf, peak, peak_src):
    """The other BASELINE.json configs, each AS STATED (same CUDA-event protocol, >= 150 ms of warm-up and of timed work):
      configs[1] in the other element type (Float64 when the headline ran Float32), with its own roofline block;
      configs[2] 2-D cdf97 lifting 4096^2 Float32 L=8: ONE image (perf/bm_dwt2_ls.jl) and a batch of 64 (fills the device);
      configs[3] full packet tree sym8 N=2^16, batch 4096;
      configs[4] 3-D db6 512^3 at L=3 (the reference's own 3-D benchmarks, benchmark/benchmarks.jl:83) and L=9 (full), and the
                 batched 2-D 4096^2 x 1024 images (cdf97 lifting and db4 filter bank, L=8) sharded 1024/G images per rank --
                 that leg also runs under torchrun (every rank its shard, max-over-ranks time) so the scaling run carries it;
      SURVEY 8f rows: 1-D cdf97 lifting, MODWT, denoise.
    Fractions are of the measured HBM peak with the compulsory byte model (2*sizeof(T) per sample per direction)."""
    import torch

    def timed_pair(fwd, inv, min_ms=150.0, sync_ranks=False):
        """>= 3 warm-up pairs and >= min_ms of warm-up work (the SM clock needs tens of ms under load to settle), then
        enough pairs for >= min_ms of timed work (at least 5)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            inv(fwd())
        e1.record(stream)
        torch.cuda.synchronize(dev)
        per = max(e0.elapsed_time(e1) / 3, 1e-3)
        if sync_ranks and world > 1:
            per = allreduce_max(per, dev)                          # every rank runs the same number of iterations
        for _ in range(int(min(200, max(0, min_ms / per - 3)))):
            inv(fwd())
        k = int(min(400, max(5, min_ms / per)))
        if sync_ranks:
            barrier()
        torch.cuda.synchronize(dev)
        e0.record(stream)
        for _ in range(k):
            inv(fwd())
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / k
        if sync_ranks and world > 1:
            ms = allreduce_max(ms, dev)
        return ms

    def entry(samples, esz, ms, **kw):
        gbs = 4.0 * esz * samples / (ms * 1e-3) / 1e9
        d = {"msamples_per_s_pair": samples / (ms * 1e-3) / 1e6, "ms_per_pair": ms, "achieved_gbs_pair": gbs,
             "frac_of_hbm_peak": gbs / peak}
        d.update(kw)
        return d

    def cleanup():
        torch.cuda.empty_cache()
        wb.release_scratch()

    res = {}
    wl = wb.wavelet(wb.WT.cdf97, wb.WT.Lifting)
    wf = wb.wavelet(wb.WT.db4)
    n2 = 4096

    # ---- configs[4], second half: 4096^2 x 1024 images sharded over the ranks (1024/G per rank; one GPU alone runs 128) ----
    per_rank = 1024 // world if world > 1 else 128
    x2 = torch.empty((per_rank, n2, n2), dtype=torch.float32, device=dev)
    for b0 in range(0, per_rank, 16):
        x2[b0:b0 + 16].normal_()
    x2 = x2.permute(2, 1, 0)                                       # column-major (n, n, B)
    for name, w in (("cdf97_lifting", wl), ("db4_filter", wf)):
        ms = timed_pair(lambda: wb.dwtc(x2, w, 8), lambda y: wb.idwtc(y, w, 8), sync_ranks=True)
        e = entry(n2 * n2 * per_rank * world, 4, ms, images_per_gpu=per_rank, n_gpus=world,
                  note="BASELINE configs[4]: 4096^2 x 1024 images split over the ranks (batch split, no data-path collective); "
                       "aggregate over all ranks / max-over-ranks device time" if world > 1 else
                       "one GPU: the 128-image shard an 8-GPU run gives each rank")
        e["frac_of_hbm_peak"] = e["achieved_gbs_pair"] / (peak * world)
        res[f"dwt2_{name}_4096x4096_f32_L8_sharded"] = e
        cleanup()
    del x2
    cleanup()
    if world > 1:
        return res

    # ---- configs[2]: 2-D cdf97 lifting, 4096^2 Float32, L = 8 -- one image (as written) and 64 images ----
    for Bi in (1, 64):
        xi = torch.randn((Bi, n2, n2), dtype=torch.float32, device=dev).permute(2, 1, 0)
        key = "dwt2_cdf97_lifting_4096x4096_f32_L8" + ("_single_image" if Bi == 1 else "")
        res[key] = entry(n2 * n2 * Bi, 4, timed_pair(lambda: wb.dwtc(xi, wl, 8), lambda y: wb.idwtc(y, wl, 8)), images=Bi,
                         **({"note": "ONE 64 MiB image: in + out fit the 126 MB L2, and a level has 2048 tiles for 148 SMs"} if Bi == 1 else {}))
        key = "dwt2_db4_filter_4096x4096_f32_L8" + ("_single_image" if Bi == 1 else "")
        res[key] = entry(n2 * n2 * Bi, 4, timed_pair(lambda: wb.dwtc(xi, wf, 8), lambda y: wb.idwtc(y, wf, 8)), images=Bi)
        del xi
        cleanup()

    # ---- configs[4], first half: 3-D db6 512^3 Float32 at L = 3 and L = 9 ----
    x3 = torch.randn((512, 512, 512), dtype=torch.float32, device=dev).permute(2, 1, 0)
    w6 = wb.wavelet(wb.WT.db6)
    for L3 in (3, 9):
        res[f"dwt3_db6_512cubed_f32_L{L3}"] = entry(512 ** 3, 4, timed_pair(lambda: wb.dwt(x3, w6, L3), lambda y: wb.idwt(y, w6, L3)), levels=L3)
    # north_star's lifting leg in 3-D (the reference has no 3-D lifting benchmark of its own): cdf97 lifting on the same cube
    res["dwt3_cdf97_lifting_512cubed_f32_L3"] = entry(512 ** 3, 4, timed_pair(lambda: wb.dwt(x3, wl, 3), lambda y: wb.idwt(y, wl, 3)), levels=3,
                                                      note="two passes per level (walk along dim 3 + the 2-D level kernel): compulsory-byte fraction is capped at 0.5")
    del x3
    cleanup()

    # ---- configs[3]: full packet tree, sym8, N = 2^16, batch 4096 ----
    Bp = 4096
    xp = torch.randn((Bp, 1 << 16), dtype=torch.float32, device=dev).t()
    w8 = wb.wavelet(wb.WT.sym8)
    ms = timed_pair(lambda: wb.wpt(xp, w8), lambda y: wb.iwpt(y, w8))
    flops_direct = 2.0 * 2 * 16 * 16 * (1 << 16) * Bp              # pair: 2 directions x 16 levels x 16 taps x 2 flop per sample
    # what the default (fast) mode executes: 12 levels in direct form + the last four levels of every 16-sample node as one
    # 16 x 16 map (16 multiply-adds per sample instead of 64); strict mode runs all 16 levels in the reference order
    flops_exec = 2.0 * 2 * (12 * 16 + 16) * (1 << 16) * Bp
    res["wpt_sym8_fulltree_65536_f32"] = entry((1 << 16) * Bp, 4, ms, signals=Bp, ms_per_1024_signals=ms * 1024 / Bp,
                                               fp32_tflops_executed=flops_exec / (ms * 1e-3) / 1e12,
                                               fp32_tflops_direct_form_equivalent=flops_direct / (ms * 1e-3) / 1e12,
                                               note="16 levels x 16 taps = 512 flop per sample per direction in direct form: FP32-pipe bound, "
                                                    "not HBM bound; three launches per direction (fused top 4 levels, on-chip subtree of 12)")
    del xp
    cleanup()

    # ---- configs[1] in the other element type, with its own roofline block ----
    other = "f64" if args.dtype == "f32" else "f32"
    oesz = 8 if other == "f64" else 4
    ocode = _lib.F64 if other == "f64" else _lib.F32
    Bo = columns_that_fit(L, _lib, dev, oesz, ocode)
    mo = measure_1d(L, _lib, dev, stream, qmf, other, Bo, max(3, min(args.steps, 6)), 3, barrier, 4242)
    mo.pop("x"); mo.pop("y")
    steps_o = max(3, min(args.steps, 6))
    gbs = 4.0 * oesz * N_SIGNAL * Bo / (mo["ms_step"] * 1e-3) / 1e9
    res[f"dwt1_db4_filter_1048576_{other}_L20"] = {
        "msamples_per_s_pair": N_SIGNAL * Bo / (mo["ms_step"] * 1e-3) / 1e6, "ms_per_pair": mo["ms_step"], "achieved_gbs_pair": gbs,
        "frac_of_hbm_peak": gbs / peak, "columns": Bo, "steps": steps_o,
        "roofline": roofline_block(mo["kern"], mo["ms_total"], steps_o, oesz, Bo, other, peak, peak_src)}
    mo.clear()
    cleanup()

    # ---- north_star's lifting leg in 1-D: cdf97 lifting, N = 2^20, L = 20, a batch of columns (in place upstream) ----
    Bl = 2048
    xl = torch.empty((Bl, N_SIGNAL), dtype=torch.float32, device=dev)
    for b0 in range(0, Bl, 256):
        xl[b0:b0 + 256].normal_()
    xl = xl.t()
    res["dwt1_cdf97_lifting_1048576_f32_L20"] = entry(N_SIGNAL * Bl, 4, timed_pair(lambda: wb.dwtc(xl, wl), lambda y: wb.idwtc(y, wl)), columns=Bl)
    del xl
    cleanup()

    # MODWT / IMODWT, db4, n = 2^20 x 64 signals, L = 10.  Compulsory bytes per direction:
    # read n + write n (L + 1) elements forward, the reverse inverse.
    nm, Bm, Lm = 1 << 20, 64, 10
    xm = torch.randn((Bm, nm), dtype=torch.float32, device=dev).t()
    ms = timed_pair(lambda: wb.modwt(xm, wf, Lm), lambda W: wb.imodwt(W, wf))
    gbs = 2.0 * (Lm + 2) * nm * Bm * 4 / (ms * 1e-3) / 1e9
    res["modwt_db4_1048576_f32_L10"] = {"msamples_per_s_pair": nm * Bm / (ms * 1e-3) / 1e6, "ms_per_pair": ms, "achieved_gbs_pair": gbs,
                                        "frac_of_hbm_peak": gbs / peak, "signals": Bm,
                                        "note": "bytes = 2 (L + 2) n B sizeof(T): the transform is (L + 1)-fold redundant"}
    del xm
    cleanup()
    # denoise = noisest (level-1 dwt + device MAD) -> dwt -> threshold -> idwt, one 1-D signal of 2^24
    # samples, sym5, L = 6, hard VisuShrink, no cycle spinning; a pure enqueue (the noise level never visits the host).
    # Compulsory bytes: read x + write y.
    nd_ = 1 << 24
    xd = torch.randn(nd_, dtype=torch.float32, device=dev)
    ms = timed_pair(lambda: wb.denoise(xd), lambda y: y)
    gbs = 2.0 * nd_ * 4 / (ms * 1e-3) / 1e9
    res["denoise_sym5_16777216_f32_L6"] = {"msamples_per_s": nd_ / (ms * 1e-3) / 1e6, "ms_per_call": ms, "ms_per_pair": ms,
                                           "achieved_gbs_pair": gbs, "frac_of_hbm_peak": gbs / peak,
                                           "note": "bytes = 2 n sizeof(T) (read x, write y); the pipeline itself makes four passes"}
    del xd
    # the same with translation-invariant cycle spinning on an image: 1024^2, 8 x 8 = 64 shifted copies run as one batch
    xi = torch.randn((1024, 1024), dtype=torch.float32, device=dev)
    ms = timed_pair(lambda: wb.denoise(xi, TI=True), lambda y: y)
    res["denoise_TI_sym5_1024x1024_f32_64spins"] = {"msamples_per_s": 1024 * 1024 / (ms * 1e-3) / 1e6, "ms_per_call": ms, "ms_per_pair": ms,
                                                    "achieved_gbs_pair": 2.0 * 1024 * 1024 * 4 / (ms * 1e-3) / 1e9,
                                                    "frac_of_hbm_peak": 2.0 * 1024 * 1024 * 4 / (ms * 1e-3) / 1e9 / peak,
                                                    "spin_msamples_per_s": 64 * 1024 * 1024 / (ms * 1e-3) / 1e6,
                                                    "note": "64 shifted dwt / threshold / idwt triples per call"}
    cleanup()
    return res


if __name__ == "__main__":
    main()
