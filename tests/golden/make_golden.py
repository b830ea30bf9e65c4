"""Generate golden vectors from the UNMODIFIED reference (PyTDSCF 1.3.3, NumPy backend).

Run in the build container only (needs /root/reference; see oracle/refshim/README.md):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Each ``<case>.npz`` holds the INPUTS a from-scratch implementation needs (MPO cores per key, site
dimensions, bond dimension, initial Hartree-product weights, step size, run options) and the
reference's OUTPUTS (initial MPS, per-step energy / autocorrelation / norm at full precision, the Krylov
iteration trace, the final MPS).  ``kernels.npz`` holds operand/result pairs of the reference's own
kernels (_op_lcr_dot, _op_lr_dot, contract_with_site_mpo, short_iterative_lanczos/arnoldi, gauge_trf,
truncate_sigvec) on small seeded random tensors.

The committed fixtures are what ``tests/test_oracle_golden.py`` pins ``oracle/tdvp_oracle.py`` against
and what the GPU parity tests compare the CUDA path with; nothing reads /root/reference at test time.
"""
from __future__ import annotations

import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference  # noqa: E402

load_reference()

import pytdscf  # noqa: E402
from discvar import HarmonicOscillator as HO  # noqa: E402
from pytdscf import _integrator, units  # noqa: E402
from pytdscf._const_cls import const  # noqa: E402
from pytdscf._contraction import (  # noqa: E402
    contract_with_site_mpo,
    multiplyH_MPS_direct_MPO,
    multiplyK_MPS_direct_MPO,
)
from pytdscf._helper import _Debug  # noqa: E402
from pytdscf._mpo_cls import OperatorCore  # noqa: E402
from pytdscf._site_cls import SiteCoef, truncate_sigvec  # noqa: E402
from pytdscf.basis import Exciton  # noqa: E402
from pytdscf.dvr_operator_cls import (  # noqa: E402
    TensorOperator,
    construct_kinetic_mpo,
    construct_nMR_recursive,
    tensor_dict_to_mpo,
)
from pytdscf.hamiltonian_cls import TensorHamiltonian  # noqa: E402
from pytdscf.model_cls import Model  # noqa: E402
from pytdscf.properties import Properties  # noqa: E402
from pytdscf.simulator_cls import Simulator  # noqa: E402


# --------------------------------------------------------------------------------------
# instrumentation of the reference (recording only; no behaviour is changed)
# --------------------------------------------------------------------------------------
RECORD: dict = {"props": [], "trace": []}

_orig_export = Properties.export_properties


def _export_and_record(self, *a, **k):
    RECORD["props"].append(
        (
            float(self.time),
            complex(self.autocorr) if self.autocorr is not None else np.nan,
            complex(self.energy) if self.energy is not None else np.nan,
            float(self.norm) if self.norm is not None else np.nan,
        )
    )
    return _orig_export(self, *a, **k)


Properties.export_properties = _export_and_record


def _wrap_solver(name):
    orig = getattr(_integrator, name)

    def wrapped(scale, multiplyOp, psi_states, thresh):
        out = orig(scale, multiplyOp, psi_states, thresh)
        kind = 0 if isinstance(multiplyOp, multiplyH_MPS_direct_MPO) else 1
        RECORD["trace"].append((kind, int(_Debug.site_now), int(_Debug.niter_krylov[_Debug.site_now])))
        return out

    setattr(_integrator, name, wrapped)


_wrap_solver("short_iterative_lanczos")
_wrap_solver("short_iterative_arnoldi")

_orig_diag = _integrator.matrix_diagonalize_lanczos


def _diag_and_record(multiplyOp, psi_states, root=0, thresh=1.0e-09):
    out = _orig_diag(multiplyOp, psi_states, root, thresh)
    RECORD["trace"].append((0, int(_Debug.site_now), int(_Debug.niter_krylov[_Debug.site_now])))
    return out


_integrator.matrix_diagonalize_lanczos = _diag_and_record


def _reset_reference_state():
    RECORD["props"].clear()
    RECORD["trace"].clear()
    _Debug.niter_krylov.clear()
    _Debug.site_now = 0


def _skip(name):
    only = os.environ.get("GOLDEN_ONLY")
    return bool(only) and only not in name


def run_reference(name, basis, operators, *, bond_dim, hartree, dt_fs, nstep, space="hilbert",
                  integrator="lanczos", conserve_norm=True, vibstate=None, thresh_sil=1e-9, relax=None, gates=None, adaptive=None):
    """Run Simulator.propagate and dump inputs + outputs to tests/golden/<name>.npz.
    ``gates``: {site: d x d matrix or length-d diagonal} applied once per step between the half sweeps
    (Model(one_gate_to_apply=...), pytdscf/_mps_cls.py:489-490, 2314-2373)."""
    if _skip(name):
        return
    _reset_reference_state()
    gate_op = None
    if gates:
        pot = {}
        for site, U in gates.items():
            U = np.asarray(U, dtype=np.complex128)
            if U.ndim == 1:
                pot[(site,)] = TensorOperator(mpo=[U.reshape(1, -1, 1)], legs=(site,))
            else:
                pot[((site, site),)] = TensorOperator(mpo=[U.reshape(1, U.shape[0], U.shape[1], 1)], legs=(site, site))
        gate_op = TensorHamiltonian(ndof=len(basis), potential=[[pot]], kinetic=None, backend="numpy")
    model = Model(basis, operators, bond_dim=bond_dim, space=space, one_gate_to_apply=gate_op)
    if hartree is not None:
        model.init_HartreeProduct = [hartree]
    if vibstate is not None:
        model.init_weight_VIBSTATE = [vibstate]
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            sim = Simulator(name, model, backend="numpy", verbose=0)
            # initial MPS exactly as the reference allocates it
            const.set_runtype(jobname=name + "_probe", space=space, integrator=integrator,
                              conserve_norm=conserve_norm, dvr=model.basinfo.is_DVR, verbose=0)
            from pytdscf._mps_mpo import MPSCoefMPO

            init = MPSCoefMPO.alloc_random(model)
            init_cores = [np.array(s.data) for s in init.superblock_states[0]]
            if relax is None:
                akw = {} if adaptive is None else dict(adaptive=True, adaptive_Dmax=adaptive[0], adaptive_dD=adaptive[1],
                                                       adaptive_p_proj=adaptive[2], adaptive_p_svd=adaptive[3])
                ener, wf = sim.propagate(stepsize=dt_fs, maxstep=nstep, thresh_sil=thresh_sil,
                                         integrator=integrator, conserve_norm=conserve_norm,
                                         energy=(space == "hilbert"), autocorr=(space == "hilbert"),
                                         norm=(space == "hilbert"), populations=False, **akw)
            else:
                import contextlib
                import io

                with contextlib.redirect_stdout(io.StringIO()):  # the reference prints a debug line per site
                    ener, wf = sim.relax(stepsize=dt_fs, maxstep=nstep, improved=(relax == "improved"), populations=False)
        finally:
            os.chdir(cwd)
    ham = model.hamiltonian
    mpo = ham.mpo[0][0]
    out = {
        "dims": np.array([len(b) for b in basis]),
        "bond_dim": np.array(bond_dim),
        "dt_au": np.array(dt_fs / units.au_in_fs),
        "nstep": np.array(nstep),
        "space": np.array(space),
        "integrator": np.array(integrator),
        "conserve_norm": np.array(conserve_norm),
        "relax": np.array("" if relax is None else relax),
        "thresh_sil": np.array(thresh_sil),
        "coupleJ": np.array(complex(ham.coupleJ[0][0])),
        "nkeys": np.array(len(mpo.operators)),
        "props": np.array([[t, a.real, a.imag, e.real, e.imag, n] for (t, a, e, n) in RECORD["props"]]),
        "trace": np.array(RECORD["trace"], dtype=np.int64),
        "final_energy": np.array(complex(ener) if ener is not None else np.nan),
    }
    for ik, (key, cores) in enumerate(mpo.operators.items()):
        out[f"key{ik}"] = np.array(repr(key))
        for ic, c in enumerate(cores):
            out[f"key{ik}_core{ic}"] = np.asarray(c)
    for i, c in enumerate(init_cores):
        out[f"init{i}"] = c
    for i, s in enumerate(wf.ci_coef.superblock_states[0]):
        out[f"final{i}"] = np.array(s.data)
    if hartree is not None:
        for i, h in enumerate(hartree):
            out[f"hartree{i}"] = np.asarray(h, dtype=np.complex128)
    for site, U in (gates or {}).items():
        out[f"gate{site}"] = np.asarray(U, dtype=np.complex128)
    if adaptive is not None:
        out["adaptive"] = np.array(adaptive, dtype=float)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"[golden] {name}: E_final={ener!r} steps={nstep} solves={len(RECORD['trace'])}")


# --------------------------------------------------------------------------------------
# cases
# --------------------------------------------------------------------------------------
def exciton_model(nprim=8):
    """The model of tests/test_exiciton_propagate.py (3 HO-DVR modes + 2-level exciton)."""
    freqs_cm1 = [1000, 2000, 3000]
    omega2 = [(f / units.au_in_cm1) ** 2 for f in freqs_cm1]
    prim = [HO(nprim, f, units="cm-1") for f in freqs_cm1] + [Exciton(nstate=2, names=["S0", "S1"])]
    dE, J, lamb, kappa = 0.01, 0.001, 0.0001, 0.0001
    W0 = np.zeros((1, nprim, 3), dtype=np.complex128)
    W1 = np.zeros((3, nprim, 4), dtype=np.complex128)
    W2 = np.zeros((4, nprim, 3), dtype=np.complex128)
    W3 = np.zeros((3, 2, 2, 1), dtype=np.complex128)
    q1 = [np.array(ho.get_grids()) for ho in prim[:3]]
    q2 = [q * q for q in q1]
    one = [np.ones_like(q) for q in q1]
    a = prim[3].get_annihilation_matrix()
    ad = prim[3].get_creation_matrix()
    W0[0, :, 0] = one[0]
    W0[0, :, 1] = q1[0]
    W0[0, :, 2] = omega2[0] / 2 * q2[0]
    W1[0, :, 0] = J * one[1] + lamb * q1[1]
    W1[0, :, 1] = one[1]
    W1[0, :, 2] = kappa * q1[1] + omega2[1] ** 2 / 2 * q2[1]
    W1[0, :, 3] = omega2[1] / 2 * q2[1]
    W1[1, :, 0] = lamb * one[1]
    W1[1, :, 2] = kappa * one[1]
    W1[2, :, 2] = one[1]
    W1[2, :, 3] = one[1]
    W2[0, :, 2] = one[2]
    W2[1, :, 0] = dE * one[2] + kappa * q1[2] + omega2[2] / 2 * q2[2]
    W2[1, :, 1] = omega2[2] / 2 * q2[2]
    W2[1, :, 2] = lamb * q1[2]
    W2[2, :, 0] = one[2]
    W2[3, :, 1] = one[2]
    W3[0, :, :, 0] = ad @ a
    W3[1, :, :, 0] = a @ ad
    W3[2, :, :, 0] = ad + a
    potential = [[{(0, 1, 2, (3, 3)): TensorOperator(mpo=[W0, W1, W2, W3], legs=(0, 1, 2, 3, 3))}]]
    kin = []
    for idof in range(3):
        d2 = prim[idof].get_2nd_derivative_matrix_dvr() / 2
        if idof == 0:
            core = np.zeros((1, nprim, nprim, 2), dtype=np.complex128)
            core[0, :, :, 0] = d2
            core[0, :, :, 1] = np.eye(nprim)
        elif idof == 2:
            core = np.zeros((2, nprim, nprim, 1), dtype=np.complex128)
            core[0, :, :, 0] = np.eye(nprim)
            core[1, :, :, 0] = d2
        else:
            core = np.zeros((2, nprim, nprim, 2), dtype=np.complex128)
            core[0, :, :, 0] = np.eye(nprim)
            core[1, :, :, 1] = np.eye(nprim)
            core[0, :, :, 1] = d2
        kin.append(core)
    kinetic = [[{((0, 0), (1, 1), (2, 2)): TensorOperator(mpo=kin, legs=(0, 0, 1, 1, 2, 2))}]]
    ham = TensorHamiltonian(ndof=4, potential=potential, kinetic=kinetic, backend="numpy")
    hartree = [ho.get_unitary()[0].tolist() for ho in prim[:3]] + [[0.0, 1.0]]
    return prim, {"hamiltonian": ham}, hartree


def henon_heiles_model(omega, lam, f, N):
    """tests/test_henon_heiles.py construction."""
    prims = [HO(N, omega) for _ in range(f)]
    w = omega / units.au_in_cm1
    func = {}
    for idof in range(f):
        if idof == 0:
            func[(0,)] = lambda Q1: pow(w, 2) / 2 * Q1**2
            if f > 1:
                func[(0, 1)] = lambda Q1, Q2: lam * pow(w, 3 / 2) * (Q1**2 * Q2)
        elif idof == f - 1:
            func[(f - 1,)] = lambda Qf: pow(w, 2) / 2 * Qf**2 - lam * pow(w, 3 / 2) / 3 * Qf**3
        else:
            func[(idof,)] = lambda Qi: pow(w, 2) / 2 * Qi**2 - lam * pow(w, 3 / 2) / 3 * Qi**3
            func[(idof, idof + 1)] = lambda Qi, Qi1: lam * pow(w, 3 / 2) * (Qi**2 * Qi1)
    mpo = construct_nMR_recursive(prims, nMR=2, func=func, rate=0.99999999999)
    K = construct_kinetic_mpo(prims)
    vib = [[0.0, 1.0] + [0.0] * (N - 2)] + [[1.0] + [0.0] * (N - 1)] * (f - 1)
    return prims, {"potential": mpo, "kinetic": K}, vib


def h2co_model():
    """BASELINE config 1: tests/h2co.tensor -> grid MPO + HO-DVR kinetic MPO."""
    freqs = [1186.325, 1252.832, 1514.908, 1831.831, 2863.96, 2916.722]
    prims = [HO(5, f, units="cm-1") for f in freqs]
    tensor_dict = pickle.load(open("/root/reference/tests/h2co.tensor", "rb"))
    mpo = tensor_dict_to_mpo(tensor_dict, rate=0.999999999999)
    K = construct_kinetic_mpo(prims)
    return prims, {"potential": mpo, "kinetic": K}


def liouville_model():
    """3 spin-1/2 sites in Liouville space (d=4): commutator of an XX+Z chain plus a Haberkorn-like
    sink on the middle site (non-Hermitian generator -> Arnoldi, conserve_norm forced False)."""
    sx = np.array([[0, 1], [1, 0]], dtype=complex) / 2
    sz = np.array([[1, 0], [0, -1]], dtype=complex) / 2
    one = np.eye(2, dtype=complex)
    P = np.array([[1, 0], [0, 0]], dtype=complex)

    def OE(op):  # O^T (x) 1   (right multiplication)
        return np.kron(op.T, one)

    def EO(op):  # 1 (x) O     (left multiplication)
        return np.kron(one, op)

    hz, Jx, k = 2.0e-3, 1.0e-3, 5.0e-4
    terms = []
    for i in range(3):
        terms.append((hz, {i: EO(sz)}))
        terms.append((-hz, {i: OE(sz)}))
    for i in range(2):
        terms.append((Jx, {i: EO(sx), i + 1: EO(sx)}))
        terms.append((-Jx, {i: OE(sx), i + 1: OE(sx)}))
    terms.append((-0.5j * k, {1: EO(P)}))
    terms.append((-0.5j * k, {1: OE(P)}))
    from pytdscf_b200.mpo_tools import sop_to_mpo

    cores = sop_to_mpo([4, 4, 4], terms)
    basis = [Exciton(nstate=4) for _ in range(3)]
    up = np.array([[1, 0], [0, 0]], dtype=complex)
    mix = np.eye(2, dtype=complex)
    hartree = [up.reshape(-1).tolist(), mix.reshape(-1).tolist(), mix.reshape(-1).tolist()]
    return basis, {"hamiltonian": cores}, hartree


# --------------------------------------------------------------------------------------
# kernel-level fixtures
# --------------------------------------------------------------------------------------
def kernel_fixtures():
    rng = np.random.default_rng(20261018)

    def crand(*shape):
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)

    const.set_runtype(jobname="golden_kernels", verbose=0)
    out = {}
    Dl, d, Dr, wl, wr = 5, 3, 4, 3, 2
    psi = crand(Dl, d, Dr)
    L = crand(Dl, wl, Dl)
    R = crand(Dr, wr, Dr)
    Wf = crand(wl, d, d, wr)
    Wd = crand(wl, d, wr)
    L1 = crand(Dl, 1, Dl)
    R1 = crand(Dr, 1, Dr)
    W_l1 = crand(1, d, d, wr)
    W_r1 = crand(wl, d, d, 1)
    Wd_l1 = crand(1, d, wr)
    Wd_r1 = crand(wl, d, 1)
    out.update(psi=psi, L=L, R=R, Wf=Wf, Wd=Wd, L1=L1, R1=R1, W_l1=W_l1, W_r1=W_r1, Wd_l1=Wd_l1, Wd_r1=Wd_r1)

    def core(data, left=False, right=False):
        key = [0, 1, 2]
        c = OperatorCore(parent_key=key, original_key=tuple(key), psite=(0 if left else (2 if right else 1)),
                         data=data, backend="numpy")
        return c

    gap = OperatorCore(parent_key=[0, 2], original_key=(0, 2), psite=1, data=1, backend="numpy")

    class _H:
        coupleJ = [[0.0]]

    mh = multiplyH_MPS_direct_MPO([[{}]], [psi], _H())
    cases = {
        "h_343": (L, core(Wf), R), "h_333": (L, core(Wd), R),
        "h_143": (Dl, core(W_l1), R), "h_133": (Dl, core(Wd_l1), R),
        "h_341": (L, core(W_r1), Dr), "h_331": (L, core(Wd_r1), Dr),
        "h_311": (L1, 0, Dr), "h_113": (Dl, 0, R1), "h_313": (L1, 0, R1),
        "h_111": (Dl, 0, Dr),
    }
    for name, (l_, c_, r_) in cases.items():
        out[name] = np.array(mh._op_lcr_dot(l_, c_, r_, psi))

    sig = crand(Dl, Dl)
    Lk = crand(Dl, wl, Dl)
    Rk = crand(Dl, wl, Dl)
    out.update(sig=sig, Lk=Lk, Rk=Rk)
    mk = multiplyK_MPS_direct_MPO([[{}]], [sig], _H())
    out["k_33"] = np.array(mk._op_lr_dot(Lk, Rk, sig))
    out["k_13"] = np.array(mk._op_lr_dot(Dl, Rk, sig))
    out["k_31"] = np.array(mk._op_lr_dot(Lk, Dl, sig))

    # environment updates
    A = crand(Dl, d, Dr)
    out["A"] = A
    sA = SiteCoef(A, "A", 1)
    sB = SiteCoef(A, "B", 1)
    out["e_A32f"] = contract_with_site_mpo(sA, sA, L, core(Wf))
    out["e_A32d"] = contract_with_site_mpo(sA, sA, L, core(Wd))
    out["e_A31f"] = contract_with_site_mpo(sA, sA, Dl, core(W_l1))
    out["e_A31d"] = contract_with_site_mpo(sA, sA, Dl, core(Wd_l1))
    out["e_A12"] = contract_with_site_mpo(sA, sA, L, gap)
    out["e_A11"] = contract_with_site_mpo(sA, sA, Dl, d)
    out["e_B32f"] = contract_with_site_mpo(sB, sB, R, core(Wf))
    out["e_B32d"] = contract_with_site_mpo(sB, sB, R, core(Wd))
    out["e_B31f"] = contract_with_site_mpo(sB, sB, Dr, core(W_r1))
    out["e_B31d"] = contract_with_site_mpo(sB, sB, Dr, core(Wd_r1))
    out["e_B12"] = contract_with_site_mpo(sB, sB, R, gap)
    out["e_B11"] = contract_with_site_mpo(sB, sB, Dr, d)

    # gauge transformations (incl. a zero-padded rank-1 tensor: LAPACK null-space completion)
    P = crand(4, 3, 5)
    out["g_psi"] = P
    a, s = SiteCoef(P.copy(), "Psi", 0).gauge_trf("Psi2Asigma")
    out["g_A"], out["g_Asig"] = np.array(a.data), np.array(s)
    b, s = SiteCoef(P.copy(), "Psi", 0).gauge_trf("Psi2sigmaB")
    out["g_B"], out["g_Bsig"] = np.array(b.data), np.array(s)
    Z = np.zeros((4, 3, 4), dtype=complex)
    Z[0, :, 0] = crand(3)
    out["g_pad"] = Z
    a, s = SiteCoef(Z.copy(), "Psi", 0).gauge_trf("Psi2Asigma")
    out["g_padA"], out["g_padAsig"] = np.array(a.data), np.array(s)
    b, s = SiteCoef(Z.copy(), "Psi", 0).gauge_trf("Psi2sigmaB")
    out["g_padB"], out["g_padBsig"] = np.array(b.data), np.array(s)

    # bond truncation
    M = crand(6, 6) * np.logspace(0, -9, 6)[None, :]
    out["t_sig"] = M
    for tag, kw in (("a", dict(p=1e-7)), ("b", dict(p=1e-3, keepdim=True)), ("c", dict(p=1e-5, regularize=True, keepdim=True))):
        U, S, Vh = truncate_sigvec(None, M.copy(), None, **kw)
        out[f"t_{tag}_U"], out[f"t_{tag}_S"], out[f"t_{tag}_Vh"] = np.array(U), np.array(S), np.array(Vh)

    # Krylov exponentials on a dense Hermitian / non-Hermitian operator
    n = 40
    Hm = crand(n, n)
    Hm = (Hm + Hm.conj().T) / 2
    Gm = Hm + 0.3 * crand(n, n)
    x0 = crand(n)
    x0 /= np.linalg.norm(x0)
    out.update(kr_H=Hm, kr_G=Gm, kr_x0=x0)

    class _Op:
        def __init__(self, M):
            self.M = M

        def stack(self, states, extend=False):
            return np.hstack([x.ravel() for x in states])  # copy, like SplitStack.stack

        def split(self, vec, truncate=False):
            return [vec.reshape(n)]

        def dot(self, states):
            return [self.M @ states[0]]

    for tag, (fn, M, cn, hist) in {
        "sil_cn": ("short_iterative_lanczos", Hm, True, 0),
        "sil_free": ("short_iterative_lanczos", Hm, False, 0),
        "sil_warm": ("short_iterative_lanczos", Hm, True, 9),
        "sil_nonherm": ("short_iterative_lanczos", Hm + 0.05j * np.diag(np.arange(n)), False, 0),
        "sia_free": ("short_iterative_arnoldi", Gm, False, 0),
        "sia_warm": ("short_iterative_arnoldi", Gm, False, 8),
    }.items():
        const.conserve_norm = cn
        _Debug.niter_krylov.clear()
        _Debug.site_now = 0
        _Debug.niter_krylov[0] = hist
        y = getattr(_integrator, fn)(-0.05j, _Op(M), [x0.copy() * (1.0 if cn else 1.7)], 1e-9)[0]
        out[f"kr_{tag}"] = np.array(y)
        out[f"kr_{tag}_niter"] = np.array(_Debug.niter_krylov[0])
    const.conserve_norm = True
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **out)
    print(f"[golden] kernels: {len(out)} arrays")


def main():
    if not os.environ.get("GOLDEN_ONLY"):
        kernel_fixtures()
    prim, ops, hartree = exciton_model()
    run_reference("exciton_D2", prim, ops, bond_dim=2, hartree=hartree, dt_fs=0.1, nstep=20)
    prim, ops, hartree = exciton_model()
    run_reference("exciton_D6", prim, ops, bond_dim=6, hartree=hartree, dt_fs=0.1, nstep=6)
    prims, ops, vib = henon_heiles_model(2000, 1.0e-3, 2, 5)
    run_reference("henon_heiles_f2", prims, ops, bond_dim=4, hartree=None, vibstate=vib, dt_fs=0.001, nstep=3)
    prims, ops, vib = henon_heiles_model(2000, 1.0e-3, 6, 5)
    run_reference("henon_heiles_f6", prims, ops, bond_dim=8, hartree=None, vibstate=vib, dt_fs=0.05, nstep=4)
    prims, ops = h2co_model()
    run_reference("h2co_D16", prims, ops, bond_dim=16, hartree=None, dt_fs=0.1, nstep=4)
    prims, ops, vib = henon_heiles_model(2000, 1.0e-3, 4, 6)
    run_reference("relax_improved_hh4", prims, ops, bond_dim=6, hartree=None, vibstate=vib, dt_fs=0.1, nstep=4, relax="improved")
    prims, ops, vib = henon_heiles_model(2000, 1.0e-3, 4, 6)
    run_reference("relax_imag_hh4", prims, ops, bond_dim=6, hartree=None, vibstate=vib, dt_fs=0.5, nstep=4, relax="imag")
    prim, ops, hartree = exciton_model()
    th = 0.3
    rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]], dtype=complex)   # kick on the exciton site
    phase = np.exp(1j * 0.05 * np.arange(8))                                               # diagonal gate on mode 1
    run_reference("gate_exciton_D6", prim, ops, bond_dim=6, hartree=hartree, dt_fs=0.1, nstep=5, gates={3: rot, 1: phase})
    prim, ops, hartree = exciton_model()
    run_reference("adaptive_exciton", prim, ops, bond_dim=1, hartree=hartree, dt_fs=0.1, nstep=6, adaptive=(8, 2, 1.0e-4, 1.0e-7))
    prims, ops, vib = henon_heiles_model(2000, 5.0e-2, 6, 5)
    run_reference("adaptive_hh6", prims, ops, bond_dim=2, hartree=None, vibstate=vib, dt_fs=0.2, nstep=5, adaptive=(5, 3, 1.0e-6, 1.0e-7))
    basis, ops, hartree = liouville_model()
    run_reference("liouville_spin3", basis, ops, bond_dim=8, hartree=hartree, dt_fs=2.0, nstep=5,
                  space="liouville", integrator="arnoldi")


if __name__ == "__main__":
    main()
