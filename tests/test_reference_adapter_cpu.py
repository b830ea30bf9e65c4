"""INTEGRATION.md executed: the UNMODIFIED reference (imported from /root/reference, build container only) runs its own
``Simulator.propagate`` with its narrow waist routed through an engine of this package by
``pytdscf_b200.reference_adapter.install`` -- here the oracle's NumPy engine, so the test is CPU-only -- and reproduces its own
golden runs: identical Krylov traces, energies / autocorrelation / norm to 1e-12.  Both patch levels: contractions only
(H_eff / K_eff term, environment update, QR gauge shift) and integrators (the reference's Lanczos / Arnoldi replaced by
``engine.krylov_expm`` fed from the reference's multiplyOp objects, warm-up table shared with the reference).
Each case runs in its own process: the reference keeps global state and the adapter monkeypatches it."""
import json
import os
import subprocess
import sys

import pytest

from oracle.reference_loader import reference_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys, tempfile
import numpy as np
sys.path.insert(0, {root!r})
from oracle.reference_loader import load_reference
load_reference()
import pytdscf
import tests.golden.make_golden as mg          # models of the golden runs + the recording hooks
from pytdscf.model_cls import Model
from pytdscf.simulator_cls import Simulator
from oracle.oracle_engine import OracleEngine
from pytdscf_b200.reference_adapter import install

case, level = {case!r}, {level!r}
eng = OracleEngine()
calls = {{"heff": 0, "keff": 0, "env": 0, "qr": 0, "krylov": 0}}
for name, key in (("heff_apply", "heff"), ("keff_apply", "keff"), ("env_update", "env"), ("qr_shift", "qr"), ("krylov_expm", "krylov")):
    def wrap(fn, key=key):
        def inner(*a, **k):
            calls[key] += 1
            return fn(*a, **k)
        return inner
    setattr(eng, name, wrap(getattr(eng, name)))
undo = install(pytdscf, eng, level=level)
if level == "integrators":                      # the recording wrappers of make_golden were replaced: wrap the new solvers
    mg._wrap_solver("short_iterative_lanczos")
    mg._wrap_solver("short_iterative_arnoldi")
mg._reset_reference_state()
kw = dict(energy=True, autocorr=True, norm=True)
if case == "exciton_D2":
    basis, ops, hartree = mg.exciton_model()
    model = Model(basis, ops, bond_dim=2); model.init_HartreeProduct = [hartree]
    run = dict(stepsize=0.1, maxstep=6)
elif case == "henon_heiles_f6":
    basis, ops, vib = mg.henon_heiles_model(2000, 1.0e-3, 6, 5)
    model = Model(basis, ops, bond_dim=8); model.init_weight_VIBSTATE = [vib]
    run = dict(stepsize=0.05, maxstep=4)
else:
    basis, ops, hartree = mg.liouville_model()
    model = Model(basis, ops, bond_dim=8, space="liouville"); model.init_HartreeProduct = [hartree]
    run = dict(stepsize=2.0, maxstep=5, integrator="arnoldi", conserve_norm=False)
    kw = dict(energy=False, autocorr=False, norm=False)
with tempfile.TemporaryDirectory() as tmp:
    os.chdir(tmp)
    sim = Simulator(case, model, backend="numpy", verbose=0)
    ener, wf = sim.propagate(populations=False, **run, **kw)
    os.chdir({root!r})
undo()
final = [np.asarray(s.data) for s in wf.ci_coef.superblock_states[0]]
np.savez({out!r}, props=np.array([[t, np.real(a), np.imag(a), np.real(e), np.imag(e), n] for (t, a, e, n) in mg.RECORD["props"]]),
         trace=np.array(mg.RECORD["trace"], dtype=np.int64), calls=json.dumps(calls), **{{f"final{{i}}": f for i, f in enumerate(final)}})
'''


@pytest.mark.skipif(not reference_available(), reason="/root/reference is only present in the build container")
@pytest.mark.parametrize("level", ["contractions", "integrators"])
@pytest.mark.parametrize("case", ["exciton_D2", "henon_heiles_f6", "liouville_spin3"])
def test_reference_runs_on_this_packages_engine(case, level, tmp_path):
    import numpy as np

    from tests.golden_io import load_run

    out = str(tmp_path / "out.npz")
    script = WORKER.format(root=ROOT, case=case, level=level, out=out)
    env = dict(os.environ, LOGURU_LEVEL="ERROR", OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=900, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    z = dict(np.load(out))
    g = load_run(case)
    n = len(z["props"])
    calls = json.loads(str(z["calls"]))
    assert calls["env"] > 0 and calls["qr"] > 0                      # the waist really went through the engine
    if level == "integrators":
        assert calls["krylov"] == len(z["trace"]) > 0
    else:
        assert calls["heff"] > 0 and calls["keff"] > 0 and calls["krylov"] == 0
    gold_trace = [tuple(t) for t in g["trace"]]
    per_step = len(gold_trace) // g["nstep"]
    assert [tuple(t) for t in z["trace"]] == gold_trace[: per_step * n]     # identical Krylov iteration counts, solve by solve
    if case != "liouville_spin3":
        for row, ref in zip(z["props"], g["props"][:n], strict=True):
            assert abs(complex(row[1], row[2]) - complex(ref[1], ref[2])) < 1e-12
            assert abs(row[3] - ref[3]) < 1e-12 * max(1.0, abs(ref[3])) and abs(row[5] - ref[5]) < 1e-12
    if n == g["nstep"]:
        for i, f in enumerate(g["final"]):
            assert np.abs(z[f"final{i}"] - f).max() < 1e-10
