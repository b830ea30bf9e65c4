"""Shared builder for the Liouville-space observable / sub-space tests (CPU host-logic test and GPU parity test):
the 3-spin model of tests/golden/make_golden_liouville_obs.py rebuilt through the product's public API."""
from __future__ import annotations

import os

import numpy as np

from tests.golden_io import GOLDEN_DIR, load_run


def load_obs():
    return dict(np.load(os.path.join(GOLDEN_DIR, "liouville_obs.npz")))


def observables_spec():
    sx = np.array([[0, 1], [1, 0]], dtype=complex) / 2
    sz = np.array([[1, 0], [0, -1]], dtype=complex) / 2
    one = np.eye(2, dtype=complex)

    def c4(m):
        return m.reshape(1, 2, 2, 1)

    return {"sz0": {((0, 0), (1, 1), (2, 2)): [c4(sz), c4(one), c4(one)]},
            "sx0sx1": {((0, 0), (1, 1)): [c4(sx), c4(sx)]},
            "P1": {((1, 1),): [c4(np.diag([1.0, 0.0]).astype(complex))]}}


def build_model(subspace):
    import pytdscf_b200 as tb

    g = load_run("liouville_spin3")
    basis = [tb.Exciton(nstate=d) for d in g["dims"]]
    pot = {key: tb.TensorOperator(mpo=[np.asarray(c) for c in cores]) for key, cores in g["operators"].items()}
    ops = {"hamiltonian": tb.TensorHamiltonian(ndof=3, potential=[[pot]], backend="cuda")}
    if subspace is None:
        for name, spec in observables_spec().items():
            p = {key: tb.TensorOperator(mpo=cores, legs=tuple(j for ind in key for j in ind)) for key, cores in spec.items()}
            ops[name] = tb.TensorHamiltonian(ndof=3, potential=[[p]], backend="cuda")
    model = tb.Model(basis, ops, bond_dim=g["bond_dim"], space="liouville", subspace_inds=subspace)
    model.init_HartreeProduct = [[h for h in g["hartree"]]]
    return g, model


def run_case(tag: str, tmp_path, engine=None):
    """Propagate 5 steps through Simulator; ``engine`` replaces the CUDA engine (CPU host-logic test)."""
    import pytdscf_b200 as tb

    sub = None if tag == "full" else {0: (0, 1, 3), 2: (0, 1, 3)}
    g, model = build_model(sub)
    os.chdir(tmp_path)
    sim = tb.Simulator("liouville_obs_" + tag, model, backend="cuda", verbose=0)
    if engine is not None:
        sim.eng = engine
    ener, wf = sim.propagate(stepsize=g["dt_au"] * tb.units.au_in_fs, maxstep=g["nstep"], thresh_sil=g["thresh_sil"],
                             integrator="arnoldi", conserve_norm=False, energy=False, autocorr=False, norm=True,
                             populations=True, observables=(sub is None), record_trace=True)
    return sim, wf


def check_case(tag: str, sim, wf, tol_expect: float, tol_state: float):
    z = load_obs()
    n = 3
    init = [z[f"{tag}_init{i}"] for i in range(n)]
    final = [z[f"{tag}_final{i}"] for i in range(n)]
    got = wf.ci_coef.to_numpy()
    for c, r in zip(got, final, strict=True):
        assert c.shape == r.shape, (c.shape, r.shape)          # identical (projected) bond / site dimensions
    assert [tuple(x.shape) for x in init] == [tuple(s.shape) for s in final]
    assert (np.array(wf.ci_coef.trace) == z[f"{tag}_trace"]).all(), "Krylov iteration trace differs"

    def dense(cores):
        v = np.asarray(cores[0])[0]
        for c in cores[1:]:
            v = np.tensordot(v, np.asarray(c), axes=(v.ndim - 1, 0))
        return v.reshape(-1)

    assert np.abs(dense(got) - dense(final)).max() <= tol_state
    for rec, nrm in zip(sim.history, z[f"{tag}_norm"], strict=True):
        assert abs(rec["norm"] - nrm) <= tol_expect and abs(rec["pops"][0] - nrm**2) <= tol_expect
    names = [str(x) for x in z[f"{tag}_names"]]
    if names:
        ref = z[f"{tag}_expect"]
        assert len(sim.history) == len(ref)
        for rec, row in zip(sim.history, ref, strict=True):
            for name, val in zip(names, row, strict=True):
                # the reference stores the real part (wavefunction.py:90-114 returns .real)
                assert abs(rec["expectations"][name] - val.real) <= tol_expect, (name, rec["expectations"][name], val)
