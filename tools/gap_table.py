"""Where the step's time is recoverable: per kernel, measured time vs the time its own roofline bound allows.

    python tools/gap_table.py profiles/r1s_bench_final.json > profiles/r1t_gap_table.txt

Input: a bench line (bench.py writes per-kernel CUDA-event times with the bound each kernel is held to).  bound time =
measured x roofline_frac; recoverable = measured - bound.  The per-kernel pass runs the three levels on one stream, so
the times add up to more than the concurrent step."""
import json
import sys

d = json.load(open(sys.argv[1]))
rows = []
for name, k in d["kernels"].items():
    ms = k["ms_per_step"]
    frac = k.get("roofline_frac") or 0.0
    rows.append((ms - ms * frac, name, k["launches_per_step"], ms, k.get("bound", "-"), frac))
rows.sort(reverse=True)
tot = sum(r[3] for r in rows)
tot_bound = sum(r[3] * r[5] for r in rows)
print(f"# {sys.argv[1]}: {d['ms_per_step']:.3f} ms/step concurrent; per-kernel pass (one stream) sums to {tot:.3f} ms, "
      f"roofline-bound sum {tot_bound:.3f} ms ({100 * tot_bound / tot:.1f} %)")
print(f"{'kernel':38s} {'n':>3s} {'ms/step':>8s} {'bound':>7s} {'frac':>6s} {'recoverable ms':>15s} {'cum %':>6s}")
cum = 0.0
for rec, name, n, ms, bound, frac in rows:
    cum += rec
    print(f"{name:38s} {int(n):3d} {ms:8.4f} {bound:>7s} {frac:6.3f} {rec:15.4f} {100 * cum / (tot - tot_bound):6.1f}")
