"""Summarise an .ncu-rep (raw page) per kernel launch: duration, DRAM bytes, pipe utilisation, stalls."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
KEYS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__inst_executed_pipe_tensor", "tensor_inst"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_inst%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("smsp__issue_active.avg.per_cycle_active", "ipc/smsp"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wavefront%"),
    ("lts__t_bytes.sum.per_second", "L2 B/s"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=" * 100)
    print(d.get("Kernel Name", "?")[:110])
    units = dict(zip(hdr, rows[1]))
    for k, label in KEYS:
        for h in hdr:
            if h == k:
                print(f"  {label:18s} {d[h]:>16s} {units.get(h, '')}")
    stalls = []
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") or ("issue_stalled" in h and "ratio" in h):
            try:
                v = float(d[h])
            except ValueError:
                continue
            if v > 0.08:
                stalls.append((v, h.split("issue_stalled_")[1].split("_per_")[0]))
    print("  stalls(per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
    for h in hdr:
        if "tensor" in h and ("pct" in h) and "sustained_active" in h:
            print("   ", h, d[h])
