"""Print a bench.py JSON line as a table (tooling)."""
import json
import sys

l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: l.get(k) for k in ["value", "ms_per_step", "gpu_launches", "clocks"]}, l.get("e2e"))
print("roofline", l.get("roofline"))
print("cpu", l.get("cpu_baseline"))
for k, v in l.get("kernels", {}).items():
    print(f"{k:32s} {v['launches_per_step']:5.1f} {v['ms_per_step']:8.3f} ms {100 * v['share']:5.1f}%")
