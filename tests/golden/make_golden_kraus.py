"""Golden run with a one-site Kraus map applied between the half sweeps (reference Model(kraus_op=...), apply_kraus /
kraus_contract_single_site, pytdscf/_mps_cls.py:2375-2418, pytdscf/kraus.py:146-222).  Build container only.

Purified-state layout of the reference's own test (tests/test_mixedstate.py:568-680): the site that carries the Kraus map has
physical dimension d * K (system index major, ancilla index minor) and the Hamiltonian acts on it as h (x) 1_K.

    python tests/golden/make_golden_kraus.py        # writes tests/golden/kraus_spin4.npz
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import importlib  # noqa: E402

mg = importlib.import_module("tests.golden.make_golden")   # loads the reference under the shims + recording hooks

from pytdscf import units  # noqa: E402
from pytdscf._const_cls import const  # noqa: E402
from pytdscf._mps_mpo import MPSCoefMPO  # noqa: E402
from pytdscf.basis import Exciton  # noqa: E402
from pytdscf.model_cls import Model  # noqa: E402
from pytdscf.simulator_cls import Simulator  # noqa: E402

from pytdscf_b200.mpo_tools import sop_to_mpo  # noqa: E402  (only builds the input MPO cores)


def main():
    run_case(two_site=False)
    run_case(two_site=True)


def run_case(two_site):
    K, d = 3, 2
    sx = np.array([[0, 1], [1, 0]], dtype=complex) / 2
    sy = np.array([[0, -1j], [1j, 0]], dtype=complex) / 2
    sz = np.array([[1, 0], [0, -1]], dtype=complex) / 2
    eK = np.eye(K, dtype=complex)
    terms = []
    if two_site:
        # the ancilla is its own site (index 2, dimension K) right of the system site 1; H acts on sites 0, 1, 3
        dims = [2, d, K, 2]
        phys = [0, 1, 3]
        for a, b in ((0, 1), (1, 3)):
            for m in (sx, sy, sz):
                terms.append((2.0e-3, {a: m, b: m}))
        for i, h in zip(phys, [1.0e-3, 3.0e-3, 0.5e-3]):
            terms.append((h, {i: sz}))
        terms.append((0.0, {2: eK}))
    else:
        dims = [2, d * K, 2, 2]
        op = lambda s, m: np.kron(m, eK) if s == 1 else m  # noqa: E731
        for i in range(3):
            for m in (sx, sy, sz):
                terms.append((2.0e-3, {i: op(i, m), i + 1: op(i + 1, m)}))
        for i, h in enumerate([1.0e-3, 3.0e-3, -2.0e-3, 0.5e-3]):
            terms.append((h, {i: op(i, sz)}))
    cores = sop_to_mpo(dims, terms)
    gam = 0.08
    B = np.array([[[1, 0], [0, np.sqrt(1 - gam)]], [[0, np.sqrt(gam)], [0, 0]]], dtype=complex)   # amplitude damping, (k, x, d)
    basis = [Exciton(nstate=n) for n in dims]
    kkey = (1, 2) if two_site else (1,)
    model = Model(basis, {"hamiltonian": cores}, bond_dim=6, kraus_op={kkey: B})
    up, mix = [1.0, 0.0], [np.sqrt(0.5), np.sqrt(0.5)]
    if two_site:
        hartree = [mix, [0.0, 1.0], [1.0] + [0.0] * (K - 1), mix]
    else:
        site1 = np.zeros(d * K)
        site1[1 * K + 0] = 1.0          # system |1>, ancilla |0>
        hartree = [mix, site1.tolist(), up, mix]
    model.init_HartreeProduct = [hartree]
    name, dt_fs, nstep = ("kraus2_spin4" if two_site else "kraus_spin4"), 2.0, 6
    mg._reset_reference_state()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            sim = Simulator(name, model, backend="numpy", verbose=0)
            const.set_runtype(jobname=name + "_probe", dvr=model.basinfo.is_DVR, verbose=0)
            init = [np.array(s.data) for s in MPSCoefMPO.alloc_random(model).superblock_states[0]]
            ener, wf = sim.propagate(stepsize=dt_fs, maxstep=nstep, autocorr=False, energy=True, norm=True, populations=False,
                                     conserve_norm=False)
        finally:
            os.chdir(cwd)
    mpo = model.hamiltonian.mpo[0][0]
    out = {"dims": np.array(dims), "bond_dim": np.array(6), "dt_au": np.array(dt_fs / units.au_in_fs), "nstep": np.array(nstep),
           "space": np.array("hilbert"), "integrator": np.array("lanczos"), "conserve_norm": np.array(False), "relax": np.array(""),
           "thresh_sil": np.array(1e-9), "coupleJ": np.array(complex(model.hamiltonian.coupleJ[0][0])), "nkeys": np.array(len(mpo.operators)),
           "props": np.array([[t, 0.0, 0.0, e.real, e.imag, n] for (t, a, e, n) in mg.RECORD["props"]]),
           "trace": np.array(mg.RECORD["trace"], dtype=np.int64), "final_energy": np.array(complex(ener)),
           "kraus_site": np.array(kkey), "kraus_B": B, "kraus_K": np.array(K)}
    for ik, (key, cs) in enumerate(mpo.operators.items()):
        out[f"key{ik}"] = np.array(repr(key))
        for ic, c in enumerate(cs):
            out[f"key{ik}_core{ic}"] = np.asarray(c)
    for i, c in enumerate(init):
        out[f"init{i}"] = c
    for i, s in enumerate(wf.ci_coef.superblock_states[0]):
        out[f"final{i}"] = np.array(s.data)
    for i, h in enumerate(hartree):
        out[f"hartree{i}"] = np.asarray(h, dtype=np.complex128)
    rd = wf.ci_coef.get_reduced_densities((0, 2))[0]           # Kraus site: (dK, dK), or the system site (d, d)
    out["rdm_site1"] = np.array(rd)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    rho = rd if two_site else np.einsum("akbk->ab", rd.reshape(d, K, d, K))
    print(f"[golden] {name}: E_final={ener!r} norm_last={mg.RECORD['props'][-1][3]:.10f} solves={len(mg.RECORD['trace'])} "
          f"rho_sys=\n{rho}")


if __name__ == "__main__":
    main()
