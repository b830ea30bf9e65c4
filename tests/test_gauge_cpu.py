"""Gauge invariance of the canonicalisation helpers (``pytdscf_b200._mps_cuda.canonicalizeA / canonicalizeB / canonicalize``),
written after the reference's own tests/test_gauge.py:13-76: the zero-padded Hartree product of ``alloc_random`` and a chain of
random cores keep their contracted tensor through every re-canonicalisation, A sites become left-isometric and B sites
right-isometric.  Host logic with the oracle's kernels injected (the GPU counterpart of every call is pinned per kernel in
tests/test_gpu_kernels.py)."""
import numpy as np
import pytest

import pytdscf_b200 as tb
from oracle.oracle_engine import OracleEngine
from pytdscf_b200._mps_cuda import MPSCoefCuda, SiteCoef, canonicalize, canonicalizeA, canonicalizeB


def contract_all(sb):
    t = np.asarray(sb[0].data)
    for s in sb[1:]:
        t = np.tensordot(t, np.asarray(s.data), axes=(-1, 0))
    return t


def assert_A(site):
    m = np.asarray(site.data).reshape(-1, site.data.shape[2])
    assert site.gauge == "A"
    np.testing.assert_allclose(m.conj().T @ m, np.eye(m.shape[1]), atol=1e-13)


def assert_B(site):
    m = np.asarray(site.data).reshape(site.data.shape[0], -1)
    assert site.gauge == "B"
    np.testing.assert_allclose(m @ m.conj().T, np.eye(m.shape[0]), atol=1e-13)


def test_gauge_of_the_initial_mps():
    """tests/test_gauge.py:16-52: five sites of d = 4, m_aux_max = 5, the reference's weights."""
    eng = OracleEngine()
    n = 5
    mpo = [np.eye(4, dtype=complex).reshape(1, 4, 4, 1) for _ in range(n)]
    key = tuple((i, i) for i in range(n))
    ham = tb.TensorHamiltonian(ndof=n, potential=[[{key: tb.TensorOperator(mpo=mpo, legs=tuple(x for i in range(n) for x in (i, i)))}]],
                               backend="cuda")
    model = tb.Model([tb.Exciton(nstate=4) for _ in range(n)], {"hamiltonian": ham}, bond_dim=5)
    model.init_HartreeProduct = [[[1.0, 0.0, 0.0, 0.0], [1.0, 1.0, 0.0, 0.0], [1.0, 1.0, 1.0, 0.0], [1.0] * 4, [1.0] * 4]]
    sb = MPSCoefCuda.alloc_random(eng, model).sites
    assert [tuple(s.data.shape) for s in sb] == [(1, 4, 4), (4, 4, 5), (5, 4, 5), (5, 4, 4), (4, 4, 1)]     # the bond rule
    contracted = contract_all(sb)
    assert abs(np.linalg.norm(contracted) - 1.0) < 1e-13
    for s in sb[1:]:
        assert_B(s)
    canonicalizeA(eng, sb[:3])
    for s in sb[:2]:
        assert_A(s)
    for s in sb[3:]:
        assert_B(s)
    np.testing.assert_allclose(contract_all(sb), contracted, atol=1e-13)
    canonicalize(eng, sb, 3)
    for s in sb[:3]:
        assert_A(s)
    assert sb[3].gauge == "Psi"
    assert_B(sb[4])
    np.testing.assert_allclose(contract_all(sb), contracted, atol=1e-13)
    canonicalize(eng, sb, 0, incremental=True)
    assert sb[0].gauge == "Psi"
    for s in sb[1:]:
        assert_B(s)
    np.testing.assert_allclose(contract_all(sb), contracted, atol=1e-13)


@pytest.mark.parametrize("center", [0, 2, 4])
def test_gauge_of_random_cores(center):
    """tests/test_gauge.py:54-76: random cores with the reference's shapes, gauge "C", canonicalised around ``center``."""
    eng = OracleEngine()
    rng = np.random.default_rng(11)
    shapes = [(1, 4, 3), (3, 4, 5), (5, 4, 4), (4, 4, 3), (3, 4, 1)]
    sb = [SiteCoef(eng.to_device(rng.random(s) + 1j * rng.random(s)), "C", i) for i, s in enumerate(shapes)]
    contracted = contract_all(sb)
    canonicalize(eng, sb, center)
    for s in sb[:center]:
        assert_A(s)
    assert sb[center].gauge == "Psi"
    for s in sb[center + 1:]:
        assert_B(s)
    np.testing.assert_allclose(contract_all(sb), contracted, atol=1e-12)
    canonicalizeB(eng, sb)                     # all the way to the left end
    for s in sb[1:]:
        assert_B(s)
    np.testing.assert_allclose(contract_all(sb), contracted, atol=1e-12)
