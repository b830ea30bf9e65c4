"""CPU tests of the synthetic BASELINE workloads: the MPO builders give the intended operators and the oracle's
TDVP reproduces the exact dense propagation when the bond dimension is full (exactness property of the
projector-splitting integrator; the reference uses the same check in tests/test_mixedstate.py:104-236)."""
import numpy as np
import pytest
import scipy.linalg

from oracle import tdvp_oracle as orc
from pytdscf_b200 import workloads
from pytdscf_b200.mpo_tools import mpo_to_dense


def dense_of(wl):
    total = None
    for key, cores in wl.operators.items():
        sites = [k[0] if isinstance(k, tuple) else k for k in key]
        full = []
        it = iter(cores)
        for p, d in enumerate(wl.dims):
            if p in sites:
                full.append(next(it))
            elif min(sites) < p < max(sites):
                raise AssertionError("gap")
            else:
                full.append(np.eye(d).reshape(1, d, d, 1))
        m = mpo_to_dense(full)
        total = m if total is None else total + m
    return total


def oracle_for(wl):
    H = orc.MPOHamiltonian(len(wl.dims), wl.operators, wl.coupleJ)
    mps = orc.initial_mps(wl.dims, wl.bond_dim, wl.hartree, space=wl.space)
    return orc.TDVPOracle(H, mps, integrator=wl.integrator, conserve_norm=wl.conserve_norm, space=wl.space)


def dense_state(cores):
    v = cores[0][0]
    for c in cores[1:]:
        v = np.tensordot(v, c, axes=(v.ndim - 1, 0))
    return v.reshape(-1)


@pytest.mark.parametrize("wl", [
    workloads.henon_heiles(f=4, N=4, D=16, lam=1e-2, dt_fs=0.2),
    workloads.pyrazine_lvc(nmode=3, nb=4, D=8, dt_fs=0.2),
    workloads.vibronic_chain(nsite=4, N=4, D=8, dt_fs=0.2),
    workloads.radical_pair(n_left=1, n_right=1, D=16, dt=0.5, spin1_right=1),
], ids=lambda w: w.name)
def test_full_rank_tdvp_is_exact(wl, monkeypatch):
    """With exact local exponentials the sweep is exact to rounding (validates MPO builders + sweep bookkeeping).
    With the reference's own Lanczos variant (alpha from v0, successive-iterate stop rule) the same
    run carries that solver's error -- up to ~3e-6 at |H dt| ~ 1 -- which parity requires us to reproduce."""
    o = oracle_for(wl)
    Hd = dense_of(wl)
    psi0 = dense_state(o.mps)
    for _ in range(3):
        o.propagate(wl.dt_au)
    exact = scipy.linalg.expm(-1j * wl.dt_au * 3 * Hd) @ psi0
    assert np.abs(dense_state(o.mps) - exact).max() < 1e-5

    def exact_solver(scale, matvec, psi, thresh, *, last_niter=0, conserve_norm=True):
        n = psi.size
        M = np.zeros((n, n), complex)
        for i in range(n):
            e = np.zeros(n, complex)
            e[i] = 1
            M[:, i] = matvec(e.reshape(psi.shape)).reshape(-1)
        return (scipy.linalg.expm(scale * M) @ psi.reshape(-1)).reshape(psi.shape), 1

    monkeypatch.setattr(orc, "sil_reference", exact_solver)
    monkeypatch.setattr(orc, "sia_reference", exact_solver)
    o = oracle_for(wl)
    Hd = dense_of(wl)
    if wl.space == "hilbert":
        np.testing.assert_allclose(Hd, Hd.conj().T, atol=1e-12)
    psi0 = dense_state(o.mps)
    e0 = o.expectation() if wl.space == "hilbert" else None
    nstep = 3
    for _ in range(nstep):
        o.propagate(wl.dt_au)
    exact = scipy.linalg.expm(-1j * wl.dt_au * nstep * Hd) @ psi0
    got = dense_state(o.mps)
    assert np.abs(got - exact).max() < 1e-11
    if wl.space == "hilbert":
        assert abs(np.linalg.norm(got) - 1) < 1e-10
        assert abs(o.expectation() - e0) < 1e-11 * max(1.0, abs(e0))
        assert abs(e0 - np.vdot(psi0, Hd @ psi0)) < 1e-12


def test_radical_pair_physics():
    """Trace of the density matrix decays with the Haberkorn rate; the generator preserves Hermiticity."""
    wl = workloads.radical_pair(n_left=1, n_right=1, D=16, dt=0.5, spin1_right=0)
    o = oracle_for(wl)
    dims = [int(round(np.sqrt(d))) for d in wl.dims]

    def rho_of(vec):
        t = vec.reshape([x for m in dims for x in (m, m)])
        n = len(dims)
        t = t.transpose(list(range(0, 2 * n, 2)) + list(range(1, 2 * n, 2)))
        return t.reshape(np.prod(dims), np.prod(dims))

    r0 = rho_of(dense_state(o.mps))
    assert abs(np.trace(r0) - 1) < 1e-12
    for _ in range(4):
        o.propagate(wl.dt_au)
    r = rho_of(dense_state(o.mps))
    np.testing.assert_allclose(r, r.conj().T, atol=1e-9)
    assert abs(np.trace(r).real - np.exp(-1e-3 * 4 * wl.dt_au)) < 1e-7  # kS = kT -> pure exponential decay


def test_default_shapes_are_the_baseline_configs():
    w3 = workloads.pyrazine_lvc()
    assert len(w3.dims) == 25 and w3.bond_dim == 256 and max(w3.dims) == 10
    w4 = workloads.radical_pair()
    assert w4.bond_dim == 1024 and set(w4.dims) == {4, 9, 16} and w4.integrator == "arnoldi" and not w4.conserve_norm
    w5 = workloads.vibronic_chain()
    assert len(w5.dims) == 128 and w5.bond_dim == 512
    w2 = workloads.henon_heiles()
    assert len(w2.dims) == 64 and w2.bond_dim == 64
    for w in (w2, w3, w4, w5):
        assert w.model().get_ndof() == len(w.dims)
