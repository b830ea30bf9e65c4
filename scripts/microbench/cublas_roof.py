"""Practical FP64 roofs on this B200: cuBLAS DGEMM / ZGEMM through torch.matmul.

Used only to obtain the denominators DESIGN.md / bench.py quote the hand-written
DMMA ZGEMM against (MEASURED_PEAKS.json has no FP64 entry).  Not on any product path.
"""
import json
import sys

import torch


def time_matmul(dtype, n, reps=5):
    a = torch.randn(n, n, device="cuda", dtype=torch.float64).to(dtype)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64).to(dtype)
    for _ in range(2):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flops = (8.0 if dtype == torch.complex128 else 2.0) * n**3
    return best, flops / (best * 1e-3) / 1e12


def main():
    out = []
    for dtype, name in ((torch.float64, "dgemm"), (torch.complex128, "zgemm")):
        for n in (1024, 2048, 4096, 8192):
            ms, tf = time_matmul(dtype, n)
            rec = {"op": name, "n": n, "ms": round(ms, 4), "tflops": round(tf, 2)}
            print(json.dumps(rec))
            out.append(rec)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
