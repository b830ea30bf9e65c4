"""CPU tests of the host-side logic: input preparation, API validation, bookkeeping, C-ABI surface."""
import os
import re
import subprocess

import numpy as np
import pytest

import pytdscf_b200 as tb
from pytdscf_b200 import _lib
from pytdscf_b200._mps_cuda import bond_dims
from pytdscf_b200.mpo_tools import mpo_to_dense, sop_to_dense, sop_to_mpo
from oracle import tdvp_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def crand(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def test_sop_to_mpo_matches_dense_sum():
    rng = np.random.default_rng(0)
    dims = [2, 3, 2, 4, 2]
    hubX, hubY = crand(rng, 4, 4), crand(rng, 4, 4)
    terms = []
    for p in range(5):
        terms.append((rng.standard_normal(), {p: crand(rng, dims[p], dims[p])}))
    for p in (0, 1, 2, 4):  # star around site 3 -> shared suffix / prefix channels
        a, b = min(p, 3), max(p, 3)
        terms.append((rng.standard_normal() + 0.3j, {p: crand(rng, dims[p], dims[p]), 3: hubX}))
        terms.append((rng.standard_normal(), {p: crand(rng, dims[p], dims[p]), 3: hubY}))
    terms.append((0.7, {0: crand(rng, 2, 2), 2: crand(rng, 2, 2), 4: crand(rng, 2, 2)}))  # 3-body term
    cores = sop_to_mpo(dims, terms)
    np.testing.assert_allclose(mpo_to_dense(cores), sop_to_dense(dims, terms), atol=1e-12)
    assert cores[0].shape[0] == 1 and cores[-1].shape[-1] == 1
    # star terms share channels: bond right of site 0 carries start, done, 2 hub channels (+1 private)
    assert cores[0].shape[-1] <= 5


def test_bond_dimension_rule():
    dims = [8, 8, 8, 2]
    assert [bond_dims(dims, i, 4) for i in range(4)] == [(1, 4), (4, 4), (4, 2), (2, 1)]
    assert [orc.bond_dims(dims, i, 4) for i in range(4)] == [bond_dims(dims, i, 4) for i in range(4)]
    assert bond_dims([5] * 6, 2, 16) == (16, 16)
    assert bond_dims([5] * 6, 1, 1000) == (5, 25)


def test_ho_dvr_basis():
    ho = tb.HarmonicOscillator(8, 1500.0, units="cm-1")
    U = ho.get_unitary()
    np.testing.assert_allclose(U.conj().T @ U, np.eye(8), atol=1e-13)
    np.testing.assert_allclose(U.conj().T @ ho.get_pos_rep_matrix() @ U, np.diag(ho.get_grids()), atol=1e-12)
    assert (U[0].real > 0).all()  # positive quadrature weights convention
    T = -0.5 * ho.get_2nd_derivative_matrix_dvr()
    V = np.diag(0.5 * ho.omega**2 * np.array(ho.get_grids()) ** 2)
    e = np.linalg.eigvalsh(T + V)
    np.testing.assert_allclose(e[:4], ho.omega * (np.arange(4) + 0.5), rtol=1e-10)
    K = tb.construct_kinetic_mpo([ho, ho, ho])
    assert [k.shape for k in K] == [(1, 8, 8, 2), (2, 8, 8, 2), (2, 8, 8, 1)]
    dense = mpo_to_dense(K)
    ref = sop_to_dense([8, 8, 8], [(1.0, {p: T}) for p in range(3)])
    np.testing.assert_allclose(dense, ref, atol=1e-12)


def _small_model(space="hilbert"):
    rng = np.random.default_rng(1)
    basis = [tb.Exciton(nstate=4), tb.Boson(3), tb.Exciton(nstate=4)]
    W = [crand(rng, 1, 4, 4, 2), crand(rng, 2, 3, 2), crand(rng, 2, 4, 4, 1)]
    return basis, W


def test_model_builds_keys_and_calc_points():
    basis, W = _small_model()
    model = tb.Model(basis, {"hamiltonian": W, "obs": W}, bond_dim=5)
    mpo = model.hamiltonian.mpo[0][0]
    assert list(mpo.operators) == [((0, 0), (1,), (2, 2))]
    assert [len(c) for c in mpo.calc_point] == [1, 1, 1]
    c0, c1, c2 = (c[0] for c in mpo.calc_point)
    assert c0.is_left_side and not c0.is_right_side and not c0.only_diag
    assert c1.only_diag and not c1.is_left_side and not c1.is_right_side
    assert c2.is_right_side
    assert "obs" in model.observables
    w, scale, m = model.initial_core_weights()
    assert m == 5 and scale == 1.0 and [len(x) for x in w] == [4, 3, 4]
    # oracle container agrees on the per-site layout
    H = orc.MPOHamiltonian(3, mpo.operators)
    assert [[(t.is_left, t.is_right, t.diag) for t in cp] for cp in H.calc_point] == \
           [[(c.is_left_side, c.is_right_side, c.only_diag) for c in cp] for cp in mpo.calc_point]


def test_potential_kinetic_split_and_scalar_term():
    ho = [tb.HarmonicOscillator(4, 1000.0), tb.HarmonicOscillator(4, 2000.0)]
    pot = [np.ones((1, 4, 2)), np.ones((2, 4, 1))]
    model = tb.Model(ho, {"potential": pot, "kinetic": tb.construct_kinetic_mpo(ho)}, bond_dim=3)
    keys = list(model.hamiltonian.mpo[0][0].operators)
    assert keys == [((0,), (1,)), ((0, 0), (1, 1))]
    assert model.basinfo.is_DVR
    w, _, _ = model.initial_core_weights()
    assert w[0].shape == (1, 4, 1) and abs(np.linalg.norm(w[0]) - 1) < 1e-14  # FBR ground state rotated to the DVR
    ham = tb.TensorHamiltonian(ndof=2, potential=[[{(): 0.25, (0, 1): tb.TensorOperator(mpo=pot)}]])
    assert ham.coupleJ[0][0] == 0.25


def test_api_validation_errors():
    basis, W = _small_model()
    with pytest.raises(ValueError):
        tb.Model(basis, {"hamiltonian": W}, space="fock")
    with pytest.raises(ValueError):
        tb.Model(basis, {"hamiltonian": W[:2]})
    with pytest.raises(ValueError):
        tb.TensorHamiltonian(ndof=3, potential=[[{(0, 1, 2): tb.TensorOperator(mpo=W)}]])  # legs mismatch
    with pytest.raises(ValueError):
        tb.TensorHamiltonian(ndof=3, potential=[[{}]], backend="tpu")
    with pytest.raises(NotImplementedError):
        tb.TensorOperator(mpo=None)
    model = tb.Model(basis, {"hamiltonian": W})
    with pytest.raises(ValueError):
        tb.Simulator("x", model, backend="jax")
    with pytest.raises(ValueError):
        tb.Simulator("x", model, ci_type="mctdh")
    from pytdscf_b200._const_cls import RunConfig

    assert RunConfig(space="Liouville", conserve_norm=True).conserve_norm is False  # forced, _const_cls.py:219-224
    with pytest.raises(ValueError):
        RunConfig(integrator="rk4")


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    basis, W = _small_model()
    sim = tb.Simulator("nogpu", tb.Model(basis, {"hamiltonian": W}, bond_dim=2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sim.propagate(maxstep=1, write_files=False)


def test_abi_library_exports_every_declared_symbol():
    """The shared library loads (CUDA runtime only, no device needed) and exports exactly the functions
    include/tdvp_b200.h declares; the ctypes prototypes cover all of them."""
    header = open(os.path.join(ROOT, "include", "tdvp_b200.h")).read()
    declared = set(re.findall(r"\b(tdvp_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load_library()
    for name in declared:
        assert hasattr(lib, name), f"{name} missing from libtdvp_b200.so"
    assert lib.tdvp_abi_version() == 2
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (tdvp_[a-z_0-9]+)\b", out))
    assert declared <= exported


def test_abi_reports_errors_without_device():
    import ctypes as C

    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load_library()
    h = C.c_void_p()
    assert lib.tdvp_create(0, None, C.byref(h)) != 0 and not h.value
    assert lib.tdvp_last_error(None) == b"null handle"


@pytest.mark.parametrize("shape", [(1, 5), (5, 1), (3, 4), (4, 4), (1, 1)])
def test_adjoint_resolves_the_lazy_conjugate_bit(shape):
    """``Engine.svd`` hands the conjugate transpose of a wide matrix to the C ABI, which reads raw memory: the tensor must hold
    the conjugated VALUES (torch's ``conj()`` alone only sets a flag, and ``.T.contiguous()`` keeps it for 1 x n / n x 1)."""
    import torch

    from pytdscf_b200._engine import adjoint

    rng = np.random.default_rng(3)
    a = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    t = adjoint(torch.from_numpy(a))
    assert t.is_contiguous() and not t.is_conj() and tuple(t.shape) == shape[::-1]
    raw = torch.view_as_real(t).numpy()                     # what a data_ptr() reader sees
    assert np.array_equal(raw[..., 0] + 1j * raw[..., 1], a.conj().T)


def test_basis_operator_matrices_match_the_reference():
    """d x d operator matrices users build MPO cores from (host-side input preparation): Boson q^2 / p^2 with the reference's
    one-level margin and the HO first-derivative matrices, against the unmodified reference when its tree is present and
    against their defining properties always."""
    b = tb.Boson(5)
    big = tb.Boson(9)
    q, p = big.get_q_matrix(), big.get_p_matrix()
    np.testing.assert_allclose(b.get_q2_matrix(), (q @ q)[:5, :5].real, atol=1e-14)      # exact matrix elements of the full space
    np.testing.assert_allclose(b.get_p2_matrix(), (p @ p)[:5, :5].real, atol=1e-14)
    np.testing.assert_allclose(0.5 * (b.get_q2_matrix() + b.get_p2_matrix()), b.get_number_matrix() + 0.5 * np.eye(5), atol=1e-14)
    ho = tb.HarmonicOscillator(7, 1500.0, units="cm-1")
    d1 = ho.get_1st_derivative_matrix_fbr()
    np.testing.assert_allclose(d1, -d1.T, atol=0)                                       # anti-Hermitian
    d1d = ho.get_1st_derivative_matrix_dvr()
    # the reference's sign convention: [D, q] = -1 on the states the truncation does not touch (DVR position operator = diag(grids))
    comm = ho.get_unitary() @ (d1d @ np.diag(ho.get_grids()) - np.diag(ho.get_grids()) @ d1d) @ ho.get_unitary().conj().T
    np.testing.assert_allclose(comm[:6, :6], -np.eye(6), atol=1e-12)
    np.testing.assert_allclose((d1 @ d1)[:5, :5], ho.get_2nd_derivative_matrix_fbr()[:5, :5], atol=1e-12)   # D^2 = d^2/dq^2 either way
    assert tb.Model([b, ho], {"hamiltonian": [np.zeros((1, 5, 1)), np.zeros((1, 7, 1))]}, bond_dim=2).get_nprim(0, 1) == 7
    if os.path.isdir("/root/reference/pytdscf"):
        from oracle.reference_loader import load_reference

        load_reference()
        from pytdscf.basis import Boson as RefBoson
        from pytdscf.basis import HarmonicOscillator as RefHO

        np.testing.assert_allclose(b.get_q2_matrix(), RefBoson(5).get_q2_matrix(), atol=1e-14)
        np.testing.assert_allclose(b.get_p2_matrix(), RefBoson(5).get_p2_matrix(), atol=1e-14)
        rho = RefHO(7, 1500.0, units="cm-1")
        np.testing.assert_allclose(d1, rho.get_1st_derivative_matrix_fbr(), atol=1e-14)
        np.testing.assert_allclose(d1d, rho.get_1st_derivative_matrix_dvr(), atol=1e-12)


def test_tensor_operator_restores_the_dense_operator():
    """``TensorOperator.restore_from_decoposed`` (reference dvr_operator_cls.py:547-554) on the kinetic MPO: the cores contract to
    sum_i -1/2 d^2/dQ_i^2 exactly."""
    ho = [tb.HarmonicOscillator(4, 1000.0 * (i + 1)) for i in range(3)]
    dense = tb.TensorOperator(mpo=tb.construct_kinetic_mpo(ho)).restore_from_decoposed()
    d = [-0.5 * h.get_2nd_derivative_matrix_dvr() for h in ho]
    eye = np.eye(4)
    ref = (np.einsum("ab,cd,ef->abcdef", d[0], eye, eye) + np.einsum("ab,cd,ef->abcdef", eye, d[1], eye)
           + np.einsum("ab,cd,ef->abcdef", eye, eye, d[2]))
    np.testing.assert_allclose(dense, ref, atol=1e-14)


def test_kinetic_operator_forms_are_the_same_operator():
    """``construct_kinetic_operator`` (reference dvr_operator_cls.py:1135-1196): the one-key MPO form and the per-mode "sop" form
    are both sum_i -c_i/2 d^2/dQ_i^2."""
    ho = [tb.HarmonicOscillator(4, 1000.0 * (i + 1)) for i in range(3)]
    coefs = [1.0, 0.5, 2.0]
    mpo_form = tb.construct_kinetic_operator(ho, coefs)
    sop_form = tb.construct_kinetic_operator(ho, coefs, forms="sop")
    assert list(mpo_form) == [((0, 0), (1, 1), (2, 2))] and list(sop_form) == [((0, 0),), ((1, 1),), ((2, 2),)]
    dense = next(iter(mpo_form.values())).restore_from_decoposed()
    eye = np.eye(4)
    d = [op.restore_from_decoposed() for op in sop_form.values()]
    ref = (np.einsum("ab,cd,ef->abcdef", d[0], eye, eye) + np.einsum("ab,cd,ef->abcdef", eye, d[1], eye)
           + np.einsum("ab,cd,ef->abcdef", eye, eye, d[2]))
    np.testing.assert_allclose(dense, ref, atol=1e-14)
    np.testing.assert_allclose(d[1], -0.25 * ho[1].get_2nd_derivative_matrix_dvr(), atol=1e-15)
    with pytest.raises(ValueError):
        tb.construct_kinetic_operator(ho, forms="dense")
