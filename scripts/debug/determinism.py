"""Run a golden case several times on the GPU; report bitwise repeatability and deviation from the reference."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden_io import load_run  # noqa: E402
from tests.test_gpu_propagation import run_cuda  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "exciton_D6"
g = load_run(name)
finals, autos = [], []
for rep in range(4):
    with tempfile.TemporaryDirectory() as tmp:
        sim, ener, wf = run_cuda(g, tmp)
    finals.append([c.copy() for c in wf.ci_coef.to_numpy()])
    autos.append(np.array([r["autocorr"] for r in sim.history]))
    ref = g["props"][:, 1] + 1j * g["props"][:, 2]
    eref = g["props"][:, 3]
    en = np.array([r["energy"] for r in sim.history])
    print(f"rep {rep}: max|autocorr - ref| = {np.abs(autos[-1] - ref).max():.3e}  max rel dE = {np.abs((en - eref) / eref).max():.3e}"
          f"  trace_ok = {(np.array(wf.ci_coef.trace) == g['trace']).all()}")
for rep in range(1, 4):
    same = all((a == b).all() for a, b in zip(finals[0], finals[rep]))
    print(f"rep {rep} bitwise identical to rep 0: {same}; max diff {max(np.abs(a - b).max() for a, b in zip(finals[0], finals[rep])):.3e}")
