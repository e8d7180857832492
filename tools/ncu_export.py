"""Condense an .ncu-rep ON THE GPU BOX into small gzip CSVs that fit gpurun's 64 MiB return limit (tooling).

    python tools/ncu_export.py gpurun_out/x.ncu-rep [--keep-rep-below-mb 40]

Writes  x.raw.csv.gz     the raw page (one row per launch, every metric)
        x.source.csv.gz  the source page reduced to: launch index, kernel, address, SASS, stall samples,
                         not-issued samples, instructions executed, thread instructions executed
and deletes the report when it is larger than the limit.
"""
import csv
import gzip
import io
import os
import subprocess
import sys

rep = sys.argv[1]
limit_mb = 40.0
if "--keep-rep-below-mb" in sys.argv:
    limit_mb = float(sys.argv[sys.argv.index("--keep-rep-below-mb") + 1])
base = rep[:-len(".ncu-rep")]

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
with gzip.open(base + ".raw.csv.gz", "wt") as fh:
    fh.write(raw)

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
KEEP = ["Address", "Source", "Warp Stall Sampling (All Samples)", "Warp Stall Sampling (Not-issued Samples)",
        "Instructions Executed", "Thread Instructions Executed"]
out = io.StringIO()
w = csv.writer(out)
w.writerow(["launch", "kernel"] + KEEP)
launch, kernel, idx = -1, "", None
for row in csv.reader(io.StringIO(src)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        launch += 1
        kernel = row[1] if len(row) > 1 else ""
        idx = None
        continue
    if row[0] == "Address":
        idx = [row.index(k) if k in row else None for k in KEEP]
        continue
    if idx is None:
        continue
    w.writerow([launch, kernel[:160]] + [row[i] if i is not None and i < len(row) else "" for i in idx])
with gzip.open(base + ".source.csv.gz", "wt") as fh:
    fh.write(out.getvalue())
mb = os.path.getsize(rep) / 2 ** 20
print(f"{rep}: {mb:.1f} MiB, {launch + 1} launches in the source page")
if mb > limit_mb:
    os.remove(rep)
    print("report removed (over the return limit); CSVs kept")
