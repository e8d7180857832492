"""Smallest possible GPU check of the 6x6-zone cases (seconds): one line per case into gpurun_out/z6_quick.log,
flushed as it goes so that a cut-off call still reports what ran.
    gpurun -- 'timeout 60 python tools/z6_quick.py'"""
import os
import sys
import time

t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
log = open(os.path.join(ROOT, "gpurun_out", "z6_quick.log"), "a")


def say(msg):
    log.write(f"[{time.time() - t0:6.2f}s] {msg}\n")
    log.flush()
    os.fsync(log.fileno())
    print(msg, flush=True)


say("start")
import torch  # noqa: E402
say(f"torch imported, cuda={torch.cuda.is_available()}")
import test_gpu_parity as T  # noqa: E402
from helpers import FusionCase  # noqa: E402

for tag in ("G416z6_L3_B2", "G416z6_L2_B1"):
    for dtype, tol in ((torch.bfloat16, T.BF16_TOL), (torch.float32, T.FP32_TOL)):
        try:
            case = FusionCase(tag)
            out, *_ = T.run_case(case, dtype)
            err = case.check_output(out, tol, f"{tag} {dtype}")
            say(f"PASS {tag} {dtype} rel-L2 {err:.3e} (tol {tol:.0e})")
        except Exception as e:  # noqa: BLE001
            say(f"FAIL {tag} {dtype}: {type(e).__name__}: {str(e)[:300]}")
say("done")
