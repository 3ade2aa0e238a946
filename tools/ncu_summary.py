#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into a small markdown table for profiles/.
    python tools/ncu_summary.py gpurun_out/r01_fused1d_f32.ncu-rep profiles/r01_fused1d_f32.md "note" """
import csv, io, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) blocks"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu summary: {rep.split('/')[-1]}", "", note, "",
             "Captured with `ncu --set full --clock-control none --import-source on` under gpurun on one B200; "
             "durations are per launch under the profiler (cold cache, serialised) and are not bench values.", ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines += [f"## `{name[:140]}`", "", "| metric | value |", "|---|---|"]
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                lines.append(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("wrote", out)


if __name__ == "__main__":
    main()
