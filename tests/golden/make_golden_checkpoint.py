"""Checkpoint interoperability fixtures, made with the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_checkpoint.py

1. ``wf_ref_exciton_D6.pkl``: the ``dill`` dump the reference's ``Simulator.propagate`` writes after 3 steps of its
   exciton test model (bond dimension 6) -- a reference OUTPUT, read back by ``pytdscf_b200.checkpoint`` in the tests.
2. ``checkpoint.npz``: the same state as arrays, the energies of 3 MORE reference steps restarted from its own file, and
   the energies of 3 more steps of the REFERENCE restarted from a file written by
   ``pytdscf_b200.write_reference_wavefunction`` (proof that the reference's restart path accepts our files: the two
   continuations must agree to rounding)."""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import tests.golden.make_golden as mg  # noqa: E402
from pytdscf.model_cls import Model  # noqa: E402
from pytdscf.simulator_cls import Simulator  # noqa: E402


def model():
    prim, ops, hartree = mg.exciton_model()
    m = Model(prim, ops, bond_dim=6)
    m.init_HartreeProduct = [hartree]
    return m


def energies():
    return [p[2].real for p in mg.RECORD["props"]]


def main():
    from pytdscf_b200.checkpoint import read_reference_wavefunction, write_reference_wavefunction

    cwd = os.getcwd()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            mg._reset_reference_state()
            sim = Simulator("ck", model(), backend="numpy", verbose=0)
            sim.propagate(stepsize=0.1, maxstep=3, energy=True, autocorr=False, norm=False, populations=False)
            shutil.copy("wf_ck.pkl", os.path.join(HERE, "wf_ref_exciton_D6.pkl"))
            d = read_reference_wavefunction("wf_ck.pkl")
            # (a) the reference restarted from its own file
            mg._reset_reference_state()
            sim = Simulator("ck", model(), backend="numpy", verbose=0)
            sim.propagate(stepsize=0.1, maxstep=3, restart=True, loadfile_ext="", savefile_ext="_a", energy=True, autocorr=False,
                          norm=False, populations=False)
            e_own = energies()
            # (b) the reference restarted from a file written by pytdscf_b200 from the same tensors
            write_reference_wavefunction("wf_ck_ours.pkl", d["cores"], d["gauges"])
            mg._reset_reference_state()
            sim = Simulator("ck", model(), backend="numpy", verbose=0)
            sim.propagate(stepsize=0.1, maxstep=3, restart=True, loadfile_ext="_ours", savefile_ext="_b", energy=True, autocorr=False,
                          norm=False, populations=False)
            e_ours = energies()
        finally:
            os.chdir(cwd)
    print("restart from the reference's file :", e_own)
    print("restart from pytdscf_b200's file  :", e_ours)
    assert np.allclose(e_own, e_ours, rtol=0, atol=1e-15), "the reference does not reproduce its own continuation from our file"
    out["gauges"] = np.array(d["gauges"])
    for i, c in enumerate(d["cores"]):
        out[f"core{i}"] = c
    out["energies_restart_reference_file"] = np.array(e_own)
    out["energies_restart_our_file"] = np.array(e_ours)
    np.savez_compressed(os.path.join(HERE, "checkpoint.npz"), **out)


if __name__ == "__main__":
    main()
