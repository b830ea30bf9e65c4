"""Golden vectors for the MPDO hermitisation rho <- (rho + rho^dagger) / 2 (reference ``MPSCoef.hermitise`` /
``svd_conj_mpdo``, pytdscf/_mps_cls.py:2289-2312, 2454-2562) from the UNMODIFIED reference.  Build container only.

    python tests/golden/make_golden_hermitise.py      # writes tests/golden/hermitise.npz

Two inputs: (a) the 3-spin Liouville chain of make_golden.py after 5 propagated steps (a nearly Hermitian MPDO: the
truncation back to the original bonds discards only rounding), (b) the same chain at bond dimension 2 with seeded random site tensors (a
genuinely non-Hermitian MPDO: the doubled bonds are truncated for real).  Stored: the site tensors before and after."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import tests.golden.make_golden as mg  # noqa: E402  (loads the reference + recording hooks)
from pytdscf._const_cls import const  # noqa: E402
from pytdscf.model_cls import Model  # noqa: E402
from pytdscf.simulator_cls import Simulator  # noqa: E402


def propagated():
    mg._reset_reference_state()
    basis, ops, hartree = mg.liouville_model()
    model = Model(basis, {"hamiltonian": ops["hamiltonian"]}, bond_dim=8, space="liouville")
    model.init_HartreeProduct = [hartree]
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            sim = Simulator("hermitise", model, backend="numpy", verbose=0)
            _, wf = sim.propagate(stepsize=2.0, maxstep=5, integrator="arnoldi", conserve_norm=False, energy=False, autocorr=False,
                                  norm=False, populations=False, observables=False)
        finally:
            os.chdir(cwd)
    return wf.ci_coef


def random_mpdo(bond_dim: int):
    from pytdscf._mps_mpo import MPSCoefMPO

    basis, ops, hartree = mg.liouville_model()
    model = Model(basis, {"hamiltonian": ops["hamiltonian"]}, bond_dim=bond_dim, space="liouville")
    model.init_HartreeProduct = [hartree]
    mps = MPSCoefMPO.alloc_random(model)
    rng = np.random.default_rng(20261018)
    for s in mps.superblock_states[0]:
        s.data = (rng.standard_normal(s.data.shape) + 1j * rng.standard_normal(s.data.shape)) / np.sqrt(s.data.size)
    return mps


def main():
    out = {}
    for tag in ("prop", "rand"):
        # "rand": bond dimension 2 < 4 = d^2, so the doubled bonds (4) are truncated for real
        mps = propagated() if tag == "prop" else random_mpdo(2)
        assert const.space == "liouville"
        sb = mps.superblock_states[0]
        before = [np.array(s.data) for s in sb]
        mps.hermitise()
        sb = mps.superblock_states[0]
        after = [np.array(s.data) for s in sb]
        out[f"{tag}_gauges"] = np.array([s.gauge for s in sb])
        for i, (a, b) in enumerate(zip(before, after, strict=True)):
            out[f"{tag}_before{i}"] = a
            out[f"{tag}_after{i}"] = b
        print(tag, [a.shape for a in before], [b.shape for b in after], [s.gauge for s in sb])
    np.savez_compressed(os.path.join(HERE, "hermitise.npz"), **out)


if __name__ == "__main__":
    main()
