"""profiles/*_sass_mnemonics.txt and *_ptxas_resources.txt from the built library / the build logs (tooling).
usage: python tools/sass_report.py <prefix>      (writes profiles/<prefix>_sass_mnemonics.txt, _ptxas_resources.txt)"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prefix = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "cfpnet_b200", "libcfp.so")], capture_output=True, text=True).stdout
pats = [("UTC*MMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"),
        ("LDGSTS", r"\bLDGSTS"), ("HMMA", r"\bHMMA"), ("FFMA2", r"\bFFMA2"), ("SYNCS", r"\bSYNCS")]
counts, name = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = [0] * len(pats)
        continue
    if name:
        for i, (_, p) in enumerate(pats):
            if re.search(p, line):
                counts[name][i] += 1
dem = dict(zip(counts, subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()))
with open(os.path.join(ROOT, "profiles", f"{prefix}_sass_mnemonics.txt"), "w") as fh:
    fh.write("# cuobjdump -sass cfpnet_b200/libcfp.so: instruction counts per kernel of the mnemonics that prove the engine\n"
             "# (B200_PROFILING.md: tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, bulk copies (TMA engine) = UBLKCP, tensor-map TMA\n"
             "#  loads = UTMALDG, cp.async = LDGSTS, mbarrier = SYNCS; HMMA would mean legacy mma.sync - there is none)\n")
    fh.write("".join(f"{n:<9s}" for n, _ in pats) + " kernel\n")
    tot = [0] * len(pats)
    for k, v in counts.items():
        if any(v[:7]):
            fh.write("".join(f"{c:<9d}" for c in v) + " " + dem[k] + "\n")
        tot = [a + b for a, b in zip(tot, v)]
    fh.write("".join(f"{c:<9d}" for c in tot) + " TOTAL (all kernels of the library)\n")
with open(os.path.join(ROOT, "profiles", f"{prefix}_ptxas_resources.txt"), "w") as fh:
    fh.write("# nvcc -Xptxas -v of the committed sources (sm_100a): registers / spills / static shared memory per kernel instantiation\n"
             "# (dynamic shared memory is set at launch; see DESIGN.md section 4)\n")
    for log in sorted(glob.glob(os.path.join(ROOT, "cfpnet_b200", "build", "*.o.log"))):
        if "_dbg" in log:
            continue
        txt = open(log).read()
        unit = os.path.basename(log)[:-6]
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\s*\nptxas info\s*: Function properties for \S+\s*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\s*\nptxas info\s*: Used (\d+) registers(.*)", txt):
            d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fh.write(f"{unit:14s} regs {int(m.group(5)):3d}  spill st/ld {int(m.group(3)):4d}/{int(m.group(4)):<4d} {m.group(6).strip(', ')[:60]:60s} {d}\n")
print("written")
