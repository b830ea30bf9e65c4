"""Pin oracle/tdvp_oracle.py against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py; PyTDSCF 1.3.3 NumPy backend under oracle/refshim).  CPU only."""
import numpy as np
import pytest

from oracle import tdvp_oracle as orc
from tests.golden_io import RUN_CASES, load_kernels, load_run


@pytest.fixture(scope="module")
def K():
    return load_kernels()


def _core(data, left=False, right=False):
    return orc.SiteCore((0, 1, 2), 1, data, left, right)


HEFF_CASES = {
    "h_343": ("L", "Wf", "R"), "h_333": ("L", "Wd", "R"), "h_143": (None, "W_l1", "R"),
    "h_133": (None, "Wd_l1", "R"), "h_341": ("L", "W_r1", None), "h_331": ("L", "Wd_r1", None),
    "h_311": ("L1", None, None), "h_113": (None, None, "R1"), "h_313": ("L1", None, "R1"),
    "h_111": (None, None, None),
}


@pytest.mark.parametrize("name", sorted(HEFF_CASES))
def test_heff_term_matches_reference(K, name):
    l_, w_, r_ = HEFF_CASES[name]
    out = orc.heff_term(None if l_ is None else K[l_], None if w_ is None else _core(K[w_]),
                        None if r_ is None else K[r_], K["psi"])
    np.testing.assert_allclose(out, K[name], rtol=0, atol=1e-13)


def test_keff_term_matches_reference(K):
    np.testing.assert_allclose(orc.keff_term(K["Lk"], K["Rk"], K["sig"]), K["k_33"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(orc.keff_term(None, K["Rk"], K["sig"]), K["k_13"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(orc.keff_term(K["Lk"], None, K["sig"]), K["k_31"], rtol=0, atol=1e-13)


ENV_CASES = {
    "e_A32f": ("A", "L", "Wf"), "e_A32d": ("A", "L", "Wd"), "e_A31f": ("A", None, "W_l1"),
    "e_A31d": ("A", None, "Wd_l1"), "e_A12": ("A", "L", None), "e_A11": ("A", None, None),
    "e_B32f": ("B", "R", "Wf"), "e_B32d": ("B", "R", "Wd"), "e_B31f": ("B", None, "W_r1"),
    "e_B31d": ("B", None, "Wd_r1"), "e_B12": ("B", "R", None), "e_B11": ("B", None, None),
}


@pytest.mark.parametrize("name", sorted(ENV_CASES))
def test_env_update_matches_reference(K, name):
    gauge, e_, w_ = ENV_CASES[name]
    out = orc.env_update_term(gauge, K["A"], K["A"], None if e_ is None else K[e_],
                              None if w_ is None else _core(K[w_]))
    np.testing.assert_allclose(out, K[name], rtol=0, atol=1e-13)


def test_gauge_shift_matches_reference(K):
    A, s = orc.shift_qr(K["g_psi"])
    np.testing.assert_allclose(A, K["g_A"], atol=1e-14)
    np.testing.assert_allclose(s, K["g_Asig"], atol=1e-14)
    s, B = orc.shift_lq(K["g_psi"])
    np.testing.assert_allclose(B, K["g_B"], atol=1e-14)
    np.testing.assert_allclose(s, K["g_Bsig"], atol=1e-14)
    # zero-padded rank-1 tensor: LAPACK Householder null-space completion (SURVEY "hard parts")
    A, s = orc.shift_qr(K["g_pad"])
    np.testing.assert_allclose(A, K["g_padA"], atol=1e-14)
    np.testing.assert_allclose(s, K["g_padAsig"], atol=1e-14)
    s, B = orc.shift_lq(K["g_pad"])
    np.testing.assert_allclose(B, K["g_padB"], atol=1e-14)
    np.testing.assert_allclose(s, K["g_padBsig"], atol=1e-14)


@pytest.mark.parametrize("tag,kw", [("a", dict(p=1e-7)), ("b", dict(p=1e-3, keepdim=True)),
                                    ("c", dict(p=1e-5, regularize=True, keepdim=True))])
def test_truncate_bond_matches_reference(K, tag, kw):
    U, S, Vh, rank = orc.truncate_bond(None, K["t_sig"], None, **kw)
    assert U.shape == K[f"t_{tag}_U"].shape and rank >= 1
    np.testing.assert_allclose(S, K[f"t_{tag}_S"], atol=1e-14)
    np.testing.assert_allclose(U @ S @ Vh, K[f"t_{tag}_U"] @ K[f"t_{tag}_S"] @ K[f"t_{tag}_Vh"], atol=1e-13)


@pytest.mark.parametrize("tag,solver,mat,cn,hist,scale0", [
    ("sil_cn", "sil", "H", True, 0, 1.0), ("sil_free", "sil", "H", False, 0, 1.7),
    ("sil_warm", "sil", "H", True, 9, 1.0), ("sil_nonherm", "sil", "Hd", False, 0, 1.7),
    ("sia_free", "sia", "G", False, 0, 1.7), ("sia_warm", "sia", "G", False, 8, 1.7)])
def test_krylov_matches_reference(K, tag, solver, mat, cn, hist, scale0):
    n = K["kr_H"].shape[0]
    M = {"H": K["kr_H"], "G": K["kr_G"], "Hd": K["kr_H"] + 0.05j * np.diag(np.arange(n))}[mat]
    fn = orc.sil_reference if solver == "sil" else orc.sia_reference
    y, niter = fn(-0.05j, lambda v: M @ v, K["kr_x0"] * scale0, 1e-9, last_niter=hist, conserve_norm=cn)
    assert niter == int(K[f"kr_{tag}_niter"])
    np.testing.assert_allclose(y, K[f"kr_{tag}"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("name", RUN_CASES)
def test_initial_mps_matches_reference(name):
    g = load_run(name)
    if g["hartree"] is None:
        pytest.skip("initial state built from HO-DVR unitary; covered by the host-API tests")
    cores = orc.initial_mps(g["dims"], g["bond_dim"], g["hartree"], space=g["space"])
    for c, r in zip(cores, g["init"], strict=True):
        assert c.shape == r.shape
        np.testing.assert_allclose(c, r, rtol=0, atol=1e-15)


@pytest.mark.parametrize("name", RUN_CASES)
def test_propagation_matches_reference(name):
    """Full propagation: identical Krylov traces, observables and final MPS (bit-level agreement observed)."""
    g = load_run(name)
    H = orc.MPOHamiltonian(len(g["dims"]), g["operators"], g["coupleJ"])
    o = orc.TDVPOracle(H, [c.copy() for c in g["init"]], thresh=g["thresh_sil"], integrator=g["integrator"],
                       conserve_norm=g["conserve_norm"], space=g["space"])
    for step in range(g["nstep"]):
        if g["space"] == "hilbert":
            t, ar, ai, er, ei, nrm = g["props"][step]
            a, e = o.autocorr(), o.expectation()
            assert abs(a - complex(ar, ai)) <= 1e-12
            assert abs(e.real - er) <= 1e-12 * max(1.0, abs(er)) and abs(o.norm() - nrm) <= 1e-12
        o.propagate(g["dt_au"])
    trace = np.array([(0 if k == "H" else 1, s, n) for k, s, n in o.trace])
    assert trace.shape == g["trace"].shape and (trace == g["trace"]).all()
    for c, r in zip(o.mps, g["final"], strict=True):
        np.testing.assert_allclose(c, r, rtol=0, atol=1e-12)


def test_reference_known_answers():
    """The literals the reference's own tests pin (tests/test_exiciton_propagate.py:178,
    tests/test_henon_heiles.py:22) at their stated tolerance (pytest.approx, rel 1e-6)."""
    g = load_run("exciton_D2")
    assert g["final_energy"].real == pytest.approx(0.010000180312707298)
    g = load_run("henon_heiles_f2")
    assert g["final_energy"].real == pytest.approx(0.018225341011652626)


def test_reference_known_reduced_density():
    """The 2 x 2 reduced density of the exciton site that the reference's own test pins as a literal
    (tests/test_exiciton_propagate.py:179-185: last record of reduced_density.nc = the state before the 20th step,
    atol 1e-9), from the oracle's state after 19 steps."""
    g = load_run("exciton_D2")
    H = orc.MPOHamiltonian(len(g["dims"]), g["operators"], g["coupleJ"])
    o = orc.TDVPOracle(H, [c.copy() for c in g["init"]], thresh=g["thresh_sil"])
    for _ in range(19):
        o.propagate(g["dt_au"])
    psi = o.mps[0]
    for c in o.mps[1:]:
        psi = np.tensordot(psi, c, axes=(-1, 0))
    m = psi.reshape(-1, g["dims"][-1])                      # (vibrational configurations, exciton state)
    rho = m.T @ m.conj()
    literal = np.array([[1.86417721e-02 + 1.60379680e-20j, 2.87367863e-02 - 6.91095824e-02j],
                        [2.87367863e-02 + 6.91095824e-02j, 9.81358228e-01 - 7.40721885e-18j]])
    np.testing.assert_allclose(rho, literal, rtol=0, atol=1e-9)


@pytest.mark.parametrize("name", ["relax_improved_hh4", "relax_imag_hh4"])
def test_relaxation_matches_reference(name):
    """Improved relaxation (Lanczos eigen-solver per site, K step skipped) and imaginary-time relaxation."""
    g = load_run(name)
    H = orc.MPOHamiltonian(len(g["dims"]), g["operators"], g["coupleJ"])
    o = orc.TDVPOracle(H, [c.copy() for c in g["init"]], thresh=g["thresh_sil"],
                       relax="improved" if g["relax"] == "improved" else True)
    for step in range(g["nstep"]):
        t, ar, ai, er, ei, nrm = g["props"][step]
        assert abs(o.expectation().real - er) <= 1e-12 * abs(er)
        assert abs(o.norm() - nrm) <= 1e-12
        o.propagate(g["dt_au"])
    trace = np.array([(0 if k == "H" else 1, s, n) for k, s, n in o.trace])
    assert trace.shape == g["trace"].shape and (trace == g["trace"]).all()
    # null-space directions of the site tensors are rounding-noise conditioned in relaxation runs (rank-deficient
    # QR inputs), so the final state is compared as the contracted coefficient vector
    def dense(cores):
        v = cores[0][0]
        for c in cores[1:]:
            v = np.tensordot(v, c, axes=(v.ndim - 1, 0))
        return v.reshape(-1)

    np.testing.assert_allclose(dense(o.mps), dense(g["final"]), rtol=0, atol=1e-12)
