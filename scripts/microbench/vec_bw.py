"""Achieved HBM bandwidth of the Krylov vector reduction (`k_dot` through `tdvp_inner`: reads two complex128 vectors, fixed-order
two-level reduction) against the measured copy bandwidth of the pool's B200 (MEASURED_PEAKS.json: 6489 GB/s)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from pytdscf_b200._engine import Engine  # noqa: E402

HBM_PEAK = 6489.0


def main():
    eng = Engine(0)
    for logn in (18, 20, 22, 24, 25):
        n = 1 << logn
        a = torch.randn(n, dtype=torch.complex128, device="cuda")
        b = torch.randn(n, dtype=torch.complex128, device="cuda")
        eng.inner(a, b, True)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(7):
            t0 = time.perf_counter()
            eng.inner(a, b, True)        # synchronous: the scalar comes back through the pinned mailbox
            best = min(best, time.perf_counter() - t0)
        gbs = 32.0 * n / best / 1e9
        print(json.dumps({"kernel": "k_dot (tdvp_inner)", "elements": n, "bytes": 32 * n, "us": round(best * 1e6, 1),
                          "GBps": round(gbs, 1), "frac_of_hbm_copy_peak": round(gbs / HBM_PEAK, 3),
                          "note": "host-timed incl. launch + scalar read-back (~10 us)"}), flush=True)


if __name__ == "__main__":
    main()
