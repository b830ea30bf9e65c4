"""Site-segment-parallel TDVP on the GPU: P processes (one per GPU over NCCL when the box has P GPUs, otherwise all on
cuda:0 with gloo staging) against the reference's MPSCoefParallel goldens (tests/golden/par_*.npz).

What can be asserted, and why:
* the HOST logic reproduces the reference run to 4e-15 when it is fed the oracle's kernels (tests/test_site_parallel_cpu.py);
* here every DEVICE kernel call made during the parallel run is replayed on the oracle's NumPy kernels with the same
  inputs (tests/checked_engine.py) and must agree to 1e-11 (gauge-invariant comparison for QR / SVD outputs), with
  identical Krylov iteration counts;
* the observables agree with the goldens only as far as the reference's scheme is defined: it multiplies by
  pinv(boundary bond matrix) and floors singular values at 1e-4, which makes its results depend on the singular-vector
  phases LAPACK happens to return (tests/golden/gauge_phase_floor_parallel.json: re-running the reference algorithm
  with random phases on the SVD vector pairs -- an exact symmetry of the SVD -- moves the norm by 2e-4 .. 3e-3).
  Tolerance = 10 x that measured floor.  For the entangled-start Frenkel cases the reference's own distribution step
  mixes two bond gauges and its result is O(1) phase-dependent, so only the per-kernel check applies there."""
import json
import os

import numpy as np
import pytest

from tests.golden_io import GOLDEN_DIR, PAR_CASES, load_parallel
from tests.mp_util import run_ranks

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLDEN_DIR, "gauge_phase_floor_parallel.json")) as _f:
    GAUGE_FLOOR = json.load(_f)
OBSERVABLE_CASES = ("par_exciton_P2", "par_hh8_P2", "par_hh8_P4")


def _worker(rank, world, port, name, tmp, use_nccl, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank if use_nccl else 0),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    try:
        import torch

        import pytdscf_b200 as tb
        from pytdscf_b200 import parallel
        from tests.checked_engine import CheckedEngine
        from tests.test_host_sweep_cpu import _build_model

        dev = rank if use_nccl else 0
        torch.cuda.set_device(dev)
        g = load_parallel(name)
        os.chdir(tmp)
        info = parallel.init_from_env("nccl" if use_nccl else "gloo")
        model = _build_model(g)
        sim = tb.Simulator(name + "_gpu", model, backend="cuda", verbose=0, device=dev)
        sim.rank_info = info
        checked = sim.eng = CheckedEngine(sim._engine())
        sim.set_initial_mps(g["init"])
        ener, wf = sim.propagate(stepsize=g["dt_fs"], maxstep=g["nstep"], parallel_split_indices=g["split"], populations=False)
        mps = wf.ci_coef
        st = sim.eng.stats()
        q.put((rank, {"history": sim.history if rank == 0 else None, "gauges": [s.gauge for s in mps.sites],
                      "shapes": [tuple(s.data.shape) for s in mps.sites], "launches": st["launches"],
                      "op_dev": dict(checked.dev), "report": checked.report()}))
        parallel.finalize(info)
    except Exception:  # pragma: no cover
        import traceback

        q.put((rank, {"error": traceback.format_exc()}))


@pytest.mark.parametrize("name", PAR_CASES)
def test_site_parallel_gpu(name, tmp_path):
    import torch

    g = load_parallel(name)
    P = g["nranks"]
    use_nccl = torch.cuda.device_count() >= P
    port = 33000 + (os.getpid() % 2000)
    res = run_ranks(_worker, P, (port, name, str(tmp_path), use_nccl))
    for r in range(P):
        assert res[r]["launches"] > 0
        if os.environ.get("PAR_DEBUG"):
            print(f"\n[rank {r}] per-kernel deviation from the oracle inside the run\n{res[r]['report']}")
        for op, dev in res[r]["op_dev"].items():
            if op.endswith("_direct"):
                continue    # raw Q factor: unique only up to the null-space completion; the product / isometry rows count
            # regularised results (singular values floored at 1e-4 exp(-s/1e-4)) inherit the uncertainty of singular
            # vectors that belong to sigma ~ 1e-12: eps/sigma ~ 1e-4 in the vector x 1e-4 weight = 1e-8 in the product
            tol = 1e-6 if op in ("qr_shift_reg_product", "svd_truncate_product") else 1e-11
            assert dev < tol, (r, op, dev, res[r]["report"])
        assert res[r]["gauges"] == g["ranks"][r]["gauges"]
        assert res[r]["shapes"] == [s.shape for s in g["ranks"][r]["sites"]]
    hist = res[0]["history"]
    assert len(hist) == g["nstep"]
    dev = (max(abs(rec["autocorr"] - complex(row[1], row[2])) for rec, row in zip(hist, g["props"])),
           max(abs(rec["energy"] - row[3]) for rec, row in zip(hist, g["props"])),
           max(abs(rec["norm"] - row[5]) for rec, row in zip(hist, g["props"])))
    if os.environ.get("PAR_DEBUG"):
        print(f"[site-parallel gpu] {name} backend={'nccl' if use_nccl else 'gloo'} max dev autocorr/energy/norm = {dev}")
    if name in OBSERVABLE_CASES:
        fl = GAUGE_FLOOR[name]
        assert dev[1] < 10 * fl["energy_abs"], (dev, fl)
        assert dev[2] < 10 * fl["norm_abs"], (dev, fl)
        assert dev[0] < 10 * fl["norm_abs"], (dev, fl)
