"""Hot SASS lines (stall samples) of one launch from a source CSV exported by tools/ncu_export.py (tooling).
usage: ncu_hot.py x.source.csv.gz <launch index> [top N]"""
import csv, gzip, sys
from collections import defaultdict
path, launch = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = [r for r in csv.reader(gzip.open(path, 'rt')) if r and r[0] == launch]
tot_s = sum(int(r[4] or 0) for r in rows); tot_i = sum(int(r[6] or 0) for r in rows)
print(rows[0][1][:150]); print("samples", tot_s, "warp-inst", tot_i, "sass lines", len(rows))
op = defaultdict(lambda: [0, 0])
for r in rows:
    o = r[3].split()[0] if not r[3].strip().startswith('@') else r[3].split()[1]
    o = o.split('.')[0]
    op[o][0] += int(r[4] or 0); op[o][1] += int(r[6] or 0)
print("by opcode (stall%, inst%):", ", ".join(f"{k}={v[0]*100/tot_s:.1f}/{v[1]*100/tot_i:.1f}" for k, v in sorted(op.items(), key=lambda kv: -kv[1][0])[:16]))
idx = {id(r): i for i, r in enumerate(rows)}
for r in sorted(rows, key=lambda r: -int(r[4] or 0))[:top]:
    i = idx[id(r)]
    print(f"{i:5d} {int(r[4] or 0)*100/tot_s:5.1f}%  inst {int(r[6] or 0)*100/tot_i:4.1f}%  {r[3].strip()[:100]}")
