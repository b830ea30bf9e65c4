"""Latency of the bond-matrix kernels: one-sided Jacobi SVD (tdvp_svd_truncate), pseudo-inverse (general and diagonal fast
path) and the Householder QR gauge shift, at the workloads' sizes.  Host-timed (these calls synchronise)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pytdscf_b200._engine import Engine  # noqa: E402


def best(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        t = min(t, time.perf_counter() - t0)
    return t


def main():
    eng = Engine(0)
    rng = np.random.default_rng(3)
    for D in (64, 256, 512, 1024):
        X = eng.to_device(rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D)))
        Xd = eng.to_device(np.diag(rng.random(D) + 0.1).astype(complex))
        d = 8
        psi = eng.to_device(rng.standard_normal((D, d, D)) + 1j * rng.standard_normal((D, d, D)))
        row = {"D": D, "svd_truncate_ms": round(best(lambda: eng.svd_truncate(X, 1e-7, keepdim=True, regularize=True)) * 1e3, 2),
               "pinv_ms": round(best(lambda: eng.pinv(X, 1e-13)) * 1e3, 2), "pinv_diagonal_ms": round(best(lambda: eng.pinv(Xd, 1e-13)) * 1e3, 3),
               "qr_shift_ms (d=8)": round(best(lambda: eng.qr_shift("A", psi)) * 1e3, 2),
               "qr_flops_model_tflops": round((16.0 * d * D * D * D - 16.0 / 3 * D**3) / best(lambda: eng.qr_shift("A", psi)) / 1e12, 2)}
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
