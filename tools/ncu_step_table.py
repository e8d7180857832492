"""Per-launch table from a raw-page CSV exported by tools/ncu_export.py (tooling)."""
import csv, gzip, re, sys
rows = list(csv.reader(gzip.open(sys.argv[1], 'rt')))
h = rows[0]
def col(n): return h.index(n) if n in h else None
def f(r, n):
    try: return float(r[col(n)])
    except Exception: return 0.0
print(f"{'#':>3s} {'kernel':20s} {'template':44s} {'us':>7s} {'Minst':>6s} {'tens%':>5s} {'ipc':>4s} {'occ%':>5s} {'regs':>4s} {'grid':>6s} {'GB/s':>6s} {'dram%':>5s} {'lsu%':>5s} {'fma%':>5s} {'alu%':>5s}")
for i, r in enumerate(rows[2:]):
    full = r[col('Kernel Name')]
    name = re.sub(r'\(.*', '', full).replace('void ', '').replace('cfp::', '')[:20]
    m = re.search(r'<(.*)>\(', full)
    targs = (m.group(1) if m else '').replace('(int)', '').replace('(bool)', '').replace('cfp::', '').replace('__nv_bfloat16', 'bf')[:44]
    print(f"{i:3d} {name:20s} {targs:44s} {f(r,'gpu__time_duration.sum'):7.1f} {f(r,'smsp__inst_executed.sum')/1e6:6.1f} "
          f"{f(r,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):5.1f} {f(r,'smsp__issue_active.avg.per_cycle_active'):4.2f} "
          f"{f(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} {r[col('launch__registers_per_thread')]:>4s} {r[col('launch__grid_size')]:>6s} "
          f"{f(r,'dram__bytes.sum.per_second'):6.0f} {f(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
          f"{f(r,'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):5.1f} {f(r,'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):5.1f} "
          f"{f(r,'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'):5.1f}")
