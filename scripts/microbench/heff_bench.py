"""H_eff / K_eff / environment-update throughput on seeded synthetic L / W / R / psi at the BASELINE bond dimensions
("H_eff matvec TFLOP/s (% FP64 peak) at D = 256 / 1024").  One MPO term per call, CUDA events, best of 5;
flops = the algorithmic formula F_H = 8 (Dl^2 Dr d wl + Dl Dr d^2 wl wr + Dl Dr^2 d wr)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pytdscf_b200._engine import Engine  # noqa: E402

PEAK = 37.0
CASES = [("config 3 site (D=256, d=10, w=6)", 256, 10, 6), ("config 5 site (D=512, d=8, w=4)", 512, 8, 4),
         ("config 4 site (D=1024, d=4, w=8)", 1024, 4, 8), ("config 4 site (D=1024, d=16, w=8)", 1024, 16, 8)]


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    eng = Engine(0)
    g = torch.Generator(device="cuda").manual_seed(7)

    def rnd(*shape):
        return torch.randn(shape, dtype=torch.complex128, device="cuda", generator=g)

    for label, D, d, w in CASES:
        psi, L, R = rnd(D, d, D), rnd(D, w, D), rnd(D, w, D)
        core = eng.upload_core(rnd(w, d, d, w).cpu().numpy())
        sig = rnd(D, D)
        fH = 8.0 * (D * D * D * d * w + D * D * d * d * w * w + D * D * D * d * w)
        fK = 16.0 * w * D**3
        tH = timeit(lambda: eng.heff_apply([(L, core, R, 1.0)], psi))
        tK = timeit(lambda: eng.keff_apply([(L, R, 1.0)], sig))
        tE = timeit(lambda: eng.env_update("A", psi, psi, L, core))
        row = {"case": label, "D": D, "d": d, "w": w,
               "heff_ms": round(tH, 4), "heff_tflops": round(fH / tH / 1e9, 2), "heff_frac_of_peak": round(fH / tH / 1e9 / PEAK, 3),
               "keff_ms": round(tK, 4), "keff_tflops": round(fK / tK / 1e9, 2),
               "env_update_ms": round(tE, 4), "env_update_tflops": round(fH / tE / 1e9, 2)}
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
