"""Tooling: clocks per tcgen05.mma (M = 128, K = 16) by N, A-view alignment, operand stride.  usage: python tools/probe/mma_probe.py"""
import ctypes
import os
import torch

here = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(here, "libmma_probe.so"))
lib.mma_probe.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 9
out = torch.zeros(148, device="cuda")


def run(N, a_off=0, lbo_a=129 * 16, ksteps=4, reps=2048, nacc=1, b_same=0, sbo_a=128, ctas=148):
    assert lib.mma_probe(out.data_ptr(), ctas, N, a_off, lbo_a, ksteps, reps, nacc, b_same, sbo_a) == 0
    v = out[:ctas].cpu()
    return float(v.median())


print("clocks per MMA (median over 148 CTAs), chain layout LBO = 129 x 16 B, aligned A start")
for N in (16, 32, 64, 128, 256):
    print(f"  N={N:3d}: one accumulator {run(N):6.1f}   two accumulators {run(N, nacc=2):6.1f}")
print("A view start offset (bytes), N = 32 / 128")
for off in (0, 16, 64, 128, 2208):
    print(f"  a_off={off:5d}: N=32 {run(32, a_off=off):6.1f}   N=128 {run(128, a_off=off):6.1f}")
print("A stride between 8-column groups (LBO), N = 32")
for lbo in (128 * 16, 129 * 16, 156 * 16, 560 * 16, 690 * 16):
    print(f"  lbo={lbo:6d}: {run(32, lbo_a=lbo, ksteps=2):6.1f}")
print("same A / B every time (ksteps = 1), N = 32 / 128")
print(f"  N=32 {run(32, ksteps=1):6.1f}   N=128 {run(128, ksteps=1):6.1f}")
print("one CTA only (no neighbours), N = 32 / 128")
print(f"  N=32 {run(32, ctas=1):6.1f}   N=128 {run(128, ctas=1):6.1f}")
