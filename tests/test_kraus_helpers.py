"""Lindblad -> Kraus helper and ancilla trace (reference pytdscf/kraus.py:17-124, :434-470) against values of the
unmodified reference (tests/golden/kraus_helpers.npz) and against the defining properties."""
import os

import numpy as np

from pytdscf_b200 import kraus
from tests.golden_io import GOLDEN_DIR

Z = dict(np.load(os.path.join(GOLDEN_DIR, "kraus_helpers.npz")))


def test_kraus_sets_reproduce_the_reference_maps():
    for name in ("spin1", "qubit_real", "d4"):
        B = kraus.lindblad_to_kraus(list(Z[name + "_L"]), float(Z[name + "_dt"]))
        d = B.shape[1]
        assert B.shape == (int(Z[name + "_k"]), d, d)
        assert np.abs(sum(np.kron(b, b.conj()) for b in B) - Z[name + "_map"]).max() < 1e-13     # same CP map (Kraus sets are unique up to a unitary mix)
        assert np.abs(sum(b.conj().T @ b for b in B) - np.eye(d)).max() < 1e-12                   # trace preserving


def test_complex_jump_operators_and_master_equation():
    """Complex L (where the reference's own assertion fails): the Kraus map must integrate d rho/dt = D rho."""
    import scipy.linalg

    rng = np.random.default_rng(11)
    d, dt = 3, 0.4
    Ls = [(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))) * 0.3 for _ in range(2)]
    B = kraus.lindblad_to_kraus(Ls, dt)
    rho = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
    rho = rho @ rho.conj().T
    rho /= np.trace(rho)
    D = np.zeros((d * d, d * d), complex)
    for L in Ls:        # independent construction: d rho = L rho L^+ - 1/2 {L^+ L, rho}, vectorised column by column
        for i in range(d * d):
            e = np.zeros(d * d, complex)
            e[i] = 1
            r = e.reshape(d, d)
            D[:, i] += (L @ r @ L.conj().T - 0.5 * (L.conj().T @ L @ r + r @ L.conj().T @ L)).reshape(-1)
    exact = (scipy.linalg.expm(D * dt) @ rho.reshape(-1)).reshape(d, d)
    got = sum(b @ rho @ b.conj().T for b in B)
    assert np.abs(got - exact).max() < 1e-13
    assert abs(np.trace(got) - 1) < 1e-13


def test_trace_kraus_dim():
    assert np.abs(kraus.trace_kraus_dim(Z["tr_in"], 3) - Z["tr_out3"]).max() < 1e-15
    assert np.abs(kraus.trace_kraus_dim(Z["tr_in"][0], 2) - Z["tr_out2"]).max() < 1e-15
