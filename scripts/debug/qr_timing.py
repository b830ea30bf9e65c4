"""Timing experiment for the QR gauge shift (run on the GPU box): TDVP_QR_DEBUG bit 0 skips the panel's column loop,
bit 1 the T-factor recursion (results are then wrong; only the durations are of interest)."""
import os
import subprocess
import sys

import numpy as np

if len(sys.argv) > 1:
    import torch

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from pytdscf_b200._engine import Engine

    eng = Engine(0)
    rng = np.random.default_rng(0)
    for (Dl, d, Dr) in [(64, 8, 64), (256, 10, 256), (512, 8, 512)]:
        psi = eng.to_device(rng.standard_normal((Dl, d, Dr)) + 1j * rng.standard_normal((Dl, d, Dr)))
        for _ in range(3):
            eng.qr_shift("A", psi)
        eng.gemm_profile(True, reset=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.qr_shift("A", psi)
        e1.record()
        torch.cuda.synchronize()
        eng.gemm_profile(False)
        br = eng.profile_breakdown()
        pan = br.get("qr.k_qr_panel_cluster", br.get("qr.k_qr_panel"))
        print(f"dbg={os.environ.get('TDVP_QR_DEBUG', '0')} shape=({Dl},{d},{Dr}) qr_shift {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us   "
              f"panel kernel {pan['ms'] / pan['launches'] * 1e3:7.1f} us x {pan['launches'] // 20} per shift", flush=True)
else:
    if os.environ.get("QR_SWEEP_ROWS"):
        for rows in ("64", "128", "256", "384"):
            print("TDVP_QR_ROWS =", rows, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "run"], env=dict(os.environ, TDVP_QR_ROWS=rows), check=False)
    else:
        for dbg in ("0", "1", "2", "3"):
            subprocess.run([sys.executable, os.path.abspath(__file__), "run"], env=dict(os.environ, TDVP_QR_DEBUG=dbg), check=False)
