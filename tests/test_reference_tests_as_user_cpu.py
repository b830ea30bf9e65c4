"""The reference's OWN test files, run as a user who switched packages would run them.

The test modules are read from ``/root/reference/tests`` at test time (build container only; skipped elsewhere -- nothing of
them is copied into this repository) and executed with ``pytdscf`` aliased to ``pytdscf_b200``, ``discvar`` to the DVR basis of
this package, and the oracle's NumPy kernels injected where a GPU engine would be created (CPU container).  What is exercised is
therefore everything ABOVE the C ABI exactly as the reference's test drives it: basis classes, operator / Hamiltonian / model
construction, ``Simulator.propagate`` keywords, output files and ``util.read_nc`` -- with the reference's own assertions (energy
literal rel 1e-6, 2 x 2 reduced-density literal atol 1e-9, tests/test_exiciton_propagate.py:178-185) deciding pass / fail.
"""
import importlib.util
import os
import sys
import types

import pytest

REF_TESTS = "/root/reference/tests"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="the reference tree exists in the build container only")


def _alias_package(monkeypatch):
    import pytdscf_b200 as tb
    from oracle.oracle_engine import OracleEngine
    from pytdscf_b200 import basis, dvr_operator_cls, hamiltonian_cls, model_cls, simulator_cls, units, util

    monkeypatch.setitem(sys.modules, "pytdscf", tb)
    for name, mod in (("basis", basis), ("dvr_operator_cls", dvr_operator_cls), ("hamiltonian_cls", hamiltonian_cls),
                      ("model_cls", model_cls), ("simulator_cls", simulator_cls), ("units", units), ("util", util)):
        monkeypatch.setitem(sys.modules, "pytdscf." + name, mod)
    discvar = types.ModuleType("discvar")
    discvar.HarmonicOscillator = tb.HarmonicOscillator
    monkeypatch.setitem(sys.modules, "discvar", discvar)
    monkeypatch.setattr(simulator_cls.Simulator, "_engine", lambda self: OracleEngine())    # no GPU here: the oracle's kernels
    monkeypatch.setattr(sys, "dont_write_bytecode", True)                                   # the reference tree is read-only


def _load_reference_test(filename: str):
    spec = importlib.util.spec_from_file_location("reference_test_" + filename.removesuffix(".py"), os.path.join(REF_TESTS, filename))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_test_exiciton_propagate_runs_unmodified(monkeypatch, tmp_path):
    """tests/test_exiciton_propagate.py of the reference: LVC exciton model on three HO-DVR modes (D = 2, 20 steps), pinned
    energy and reduced density.  The only change a switching user makes is the backend string."""
    _alias_package(monkeypatch)
    monkeypatch.chdir(tmp_path)
    mod = _load_reference_test("test_exiciton_propagate.py")
    mod.test_exiciton_propagate(backend="cuda")
    out = sorted(os.listdir(tmp_path / "LVC_Exciton_test_prop"))
    assert {"main.log", "autocorr.dat", "populations.dat", "expectations.dat", "reduced_density.nc"} <= set(out)
    assert os.path.exists(tmp_path / "wf_LVC_Exciton_test.pkl")


def test_reference_backend_strings_are_refused(monkeypatch, tmp_path):
    """The same test with the reference's own backend strings: there is no NumPy / JAX dispatch in this package, so the
    switch is explicit (the backend string is the boundary)."""
    _alias_package(monkeypatch)
    monkeypatch.chdir(tmp_path)
    mod = _load_reference_test("test_exiciton_propagate.py")
    for backend in ("numpy", "jax"):
        with pytest.raises(ValueError, match="backend"):
            mod.test_exiciton_propagate(backend=backend)
