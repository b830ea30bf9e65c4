"""Golden vectors for Liouville-space expectation values Tr(O rho) (reference ``_exp_liouville``,
pytdscf/_mps_cls.py:3769-3838) and for the sub-space projection of the initial MPDO (``project_subspace``,
pytdscf/_mps_mpo.py:196-220), from the UNMODIFIED reference.  Build container only.

    python tests/golden/make_golden_liouville_obs.py      # writes tests/golden/liouville_obs.npz

Model: the 3-spin Liouville chain of make_golden.py (``liouville_spin3``: commutator of an XX+Z chain + sink on the middle
site, Arnoldi).  Observables (Hilbert-space MPOs, physical dimension 2): sz on site 0 as a full-length MPO of 4-index
cores, the two-site product sx0 sx1 on a key that stops before the last site, and the projector on site 1 alone.  A second run restricts sites 0 and 2 to the populations-and-one-coherence sub-space (0, 1, 3)."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import tests.golden.make_golden as mg  # noqa: E402  (loads the reference + recording hooks)
from pytdscf._const_cls import const  # noqa: E402
from pytdscf.dvr_operator_cls import TensorOperator  # noqa: E402
from pytdscf.hamiltonian_cls import TensorHamiltonian  # noqa: E402
from pytdscf.model_cls import Model  # noqa: E402
from pytdscf.properties import Properties  # noqa: E402
from pytdscf.simulator_cls import Simulator  # noqa: E402

EXPECT: list = []
_orig = Properties.export_properties


def _rec(self, *a, **k):
    out = _orig(self, *a, **k)
    EXPECT.append({key: complex(v) for key, v in self.expectations.items()})
    return out


Properties.export_properties = _rec


def observables():
    sx = np.array([[0, 1], [1, 0]], dtype=complex) / 2
    sz = np.array([[1, 0], [0, -1]], dtype=complex) / 2
    one = np.eye(2, dtype=complex)
    c4 = lambda m: m.reshape(1, 2, 2, 1)  # noqa: E731
    obs = {
        "sz0": ({((0, 0), (1, 1), (2, 2)): [c4(sz), c4(one), c4(one)]}),
        "sx0sx1": ({((0, 0), (1, 1)): [c4(sx), c4(sx)]}),
        # (a diagonal 3-index core cannot be used here: the reference stores its key entry as an int, which its
        #  _exp_liouville cannot iterate -- pytdscf/_mps_cls.py:3808)
        "P1": ({((1, 1),): [c4(np.diag([1.0, 0.0]).astype(complex))]}),
    }
    return obs


def to_ham(spec):
    pot = {}
    for key, cores in spec.items():
        legs = tuple(j for ind in key for j in (ind if isinstance(ind, tuple) else (ind,)))
        pot[key] = TensorOperator(mpo=cores, legs=legs)   # _exp_liouville needs every key entry as a tuple: (i,) or (i, i)
    return TensorHamiltonian(ndof=3, potential=[[pot]], kinetic=None, backend="numpy")


def run(subspace):
    EXPECT.clear()
    mg._reset_reference_state()
    basis, ops, hartree = mg.liouville_model()
    operators = {"hamiltonian": ops["hamiltonian"]}
    if subspace is None:
        # (with subspace_inds the reference projects every observable like the Liouvillian, hamiltonian_cls.py:852-879,
        #  which only works for cores of the Liouville dimension: Hilbert-space observables and sub-spaces exclude each other)
        for name, spec in observables().items():
            operators[name] = to_ham(spec)
    model = Model(basis, operators, bond_dim=8, space="liouville", subspace_inds=subspace)
    model.init_HartreeProduct = [hartree]
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            sim = Simulator("liouville_obs", model, backend="numpy", verbose=0)
            const.set_runtype(jobname="probe", space="liouville", integrator="arnoldi", conserve_norm=False, dvr=model.basinfo.is_DVR, verbose=0)
            from pytdscf._mps_mpo import MPSCoefMPO

            init = MPSCoefMPO.alloc_random(model)
            init_cores = [np.array(s.data) for s in init.superblock_states[0]]
            ener, wf = sim.propagate(stepsize=2.0, maxstep=5, integrator="arnoldi", conserve_norm=False, energy=False, autocorr=False,
                                     norm=True, populations=True, observables=subspace is None)
        finally:
            os.chdir(cwd)
    final = [np.array(s.data) for s in wf.ci_coef.superblock_states[0]]
    return init_cores, final, list(EXPECT), list(mg.RECORD["trace"]), [p[3] for p in mg.RECORD["props"]]


def main():
    out = {}
    for tag, sub in (("full", None), ("sub", {0: (0, 1, 3), 2: (0, 1, 3)})):
        init, final, exps, trace, norms = run(sub)
        out[f"{tag}_norm"] = np.array(norms)     # |centre tensor| (reference MPSCoef.norm, also what populations.dat holds squared)
        names = list(exps[0].keys()) if exps else []
        out[f"{tag}_names"] = np.array(names)
        out[f"{tag}_expect"] = np.array([[e[n] for n in names] for e in exps])
        out[f"{tag}_trace"] = np.array(trace, dtype=np.int64)
        for i, (a, b) in enumerate(zip(init, final, strict=True)):
            out[f"{tag}_init{i}"] = a
            out[f"{tag}_final{i}"] = b
        print(tag, names, out[f"{tag}_expect"][:2], [a.shape for a in init], len(trace))
    np.savez_compressed(os.path.join(HERE, "liouville_obs.npz"), **out)


if __name__ == "__main__":
    main()
