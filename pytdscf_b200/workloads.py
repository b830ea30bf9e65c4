"""Synthetic workloads of the BASELINE.json configurations (input preparation, host side).

Each builder returns a ``Workload``: the MPO cores per key, site dimensions, bond dimension, initial
Hartree-product weights and run options -- everything ``Simulator`` (GPU) and the CPU oracle need to run the
SAME propagation.  Parameters are seeded (``np.random.default_rng(1234 + config_index)``); the
MPO patterns follow the reference's tests / notebooks:

  c1  H2CO 6-mode grid MPO, D=16            tests/golden/h2co_D16.npz (made from the reference's tests/h2co.tensor)
  c2  Henon-Heiles, f modes, HO-DVR, D=64   tests/test_henon_heiles.py:47-73 (diag potential MPO + kinetic MPO)
  c3  pyrazine-like LVC, 24 modes + exciton, D=256   docs/notebook/pyrazine-qvc.ipynb (one full-length 4-index MPO)
  c4  radical pair in Liouville space, D=1024, Arnoldi   docs/notebook/radicalpair-liouville.ipynb
  c5  128-site vibronic chain, D=512 (site-parallel shape)   tests/test_mpi_exiciton_propagate.py:70-190 tiled
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import units
from .basis import Boson, Exciton, HarmonicOscillator
from .dvr_operator_cls import construct_kinetic_mpo
from .mpo_tools import sop_to_mpo


@dataclass
class Workload:
    name: str
    dims: list[int]
    operators: dict  # key -> list of cores
    hartree: list  # per-site 1-D weights or 3-D cores
    bond_dim: int
    dt_fs: float
    space: str = "hilbert"
    integrator: str = "lanczos"
    conserve_norm: bool = True
    coupleJ: complex = 0.0
    description: str = ""
    meta: dict = field(default_factory=dict)

    @property
    def dt_au(self) -> float:
        return self.dt_fs / units.au_in_fs

    def model(self):
        """The ``pytdscf_b200.Model`` of this workload."""
        from .dvr_operator_cls import TensorOperator
        from .hamiltonian_cls import TensorHamiltonian
        from .model_cls import Model

        basis = [Exciton(nstate=d) for d in self.dims]
        pot = {key: TensorOperator(mpo=list(cores)) for key, cores in self.operators.items()}
        if self.coupleJ != 0:
            pot[()] = self.coupleJ
        ham = TensorHamiltonian(ndof=len(basis), potential=[[pot]], backend="cuda")
        m = Model(basis, {"hamiltonian": ham}, bond_dim=self.bond_dim, space=self.space)
        m.init_HartreeProduct = [list(self.hartree)]
        return m


def diag_cores(cores: list[np.ndarray]) -> list[np.ndarray]:
    """4-index cores that are diagonal in (bra, ket) -> 3-index cores (w_l, d, w_r)."""
    out = []
    for W in cores:
        d = W.shape[1]
        idx = np.arange(d)
        D = W[:, idx, idx, :]
        chk = W.copy()
        chk[:, idx, idx, :] = 0
        if np.abs(chk).max() > 0:
            raise ValueError("core is not diagonal")
        out.append(np.ascontiguousarray(D))
    return out


def _key_full(n):
    return tuple((i, i) for i in range(n))


def _key_diag(n):
    return tuple((i,) for i in range(n))


# ------------------------------------------------------------------------------------------------
def henon_heiles(f: int = 64, N: int = 8, D: int = 64, omega_cm1: float = 2000.0, lam: float = 1.0e-3,
                 dt_fs: float = 0.05) -> Workload:
    """c2: H = 1/2 sum(-d^2/dQ^2 + w^2 Q^2) + lam w^1.5 sum(Q_i^2 Q_{i+1} - Q_{i+1}^3/3) on HO-DVR grids."""
    prims = [HarmonicOscillator(N, omega_cm1, units="cm-1") for _ in range(f)]
    w = omega_cm1 / units.au_in_cm1
    terms = []
    for i, p in enumerate(prims):
        q = np.array(p.get_grids())
        v = 0.5 * w**2 * q**2
        if i > 0:
            v = v - lam * w**1.5 / 3.0 * q**3
        terms.append((1.0, {i: np.diag(v)}))
        if i < f - 1:
            qn = np.array(prims[i + 1].get_grids())
            terms.append((lam * w**1.5, {i: np.diag(q**2), i + 1: np.diag(qn)}))
    pot = diag_cores(sop_to_mpo([N] * f, terms))
    kin = construct_kinetic_mpo(prims)
    # first mode vibrationally excited, others in the ground state (tests/test_henon_heiles.py:74-78)
    hartree = []
    for i, p in enumerate(prims):
        v = np.zeros(N, dtype=complex)
        v[1 if i == 0 else 0] = 1.0
        hartree.append((v @ p.get_unitary()).reshape(1, N, 1))
    return Workload(f"c2_henon_heiles_f{f}_N{N}_D{D}", [N] * f, {_key_diag(f): pot, _key_full(f): kin}, hartree, D, dt_fs,
                    description=f"Henon-Heiles {f} modes, HO-DVR N={N}, diag potential MPO + kinetic MPO, D={D}")


def pyrazine_lvc(nmode: int = 24, nb: int = 10, D: int = 256, dt_fs: float = 0.05, seed: int = 1236) -> Workload:
    """c3: two-state linear vibronic coupling, exciton site (d=2) + ``nmode`` boson sites, one full-length MPO."""
    rng = np.random.default_rng(seed)
    n = nmode + 1
    dims = [2] + [nb] * nmode
    omega = rng.uniform(0.0015, 0.0145, nmode)  # ~330..3200 cm-1
    kappa = rng.normal(0.0, 0.004, (nmode, 2))
    lam = np.where(rng.random(nmode) < 0.25, rng.normal(0.0, 0.006, nmode), 0.0)
    lam[0] = 0.0095  # at least one coupling mode
    dE = 0.031  # ~0.85 eV gap
    P = [np.diag([1.0, 0.0]), np.diag([0.0, 1.0])]
    sx = np.array([[0.0, 1.0], [1.0, 0.0]])
    b = Boson(nb)
    qm, num = b.get_q_matrix(), b.get_number_matrix()
    terms = [(-dE / 2, {0: P[0]}), (dE / 2, {0: P[1]})]
    for k in range(nmode):
        terms.append((omega[k], {k + 1: num + 0.5 * np.eye(nb)}))
        for s in range(2):
            terms.append((kappa[k, s], {0: P[s], k + 1: qm}))
        if lam[k] != 0.0:
            terms.append((lam[k], {0: sx, k + 1: qm}))
    cores = sop_to_mpo(dims, terms)
    hartree = [[0.0, 1.0]] + [[1.0] + [0.0] * (nb - 1)] * nmode  # vertical excitation to S2, vibrational ground state
    return Workload(f"c3_pyrazine_lvc_{nmode}mode_D{D}", dims, {_key_full(n): cores}, hartree, D, dt_fs,
                    description=f"LVC {nmode} boson modes (d={nb}) + 2-level exciton site, full-length 4-index MPO "
                                f"w={max(c.shape[-1] for c in cores)}, D={D}", meta={"w": [c.shape[-1] for c in cores[:-1]]})


def _spin_ops(mult: int):
    s = (mult - 1) / 2
    m = np.arange(s, -s - 1, -1)
    sz = np.diag(m).astype(complex)
    sp = np.zeros((mult, mult), dtype=complex)
    for i in range(1, mult):
        sp[i - 1, i] = math.sqrt(s * (s + 1) - m[i] * (m[i] + 1))
    sx = (sp + sp.T.conj()) / 2
    sy = (sp - sp.T.conj()) / 2j
    return sx, sy, sz


def radical_pair(n_left: int = 8, n_right: int = 8, D: int = 1024, dt: float = 0.5, seed: int = 1237,
                 spin1_right: int = 1) -> Workload:
    """c4: radical pair with isotropic hyperfine couplings in Liouville space (vectorised density matrix).

    Sites: ``n_left`` nuclei | electron pair (d = 16) | ``n_right`` nuclei; spin-1/2 nuclei have d = 4, the
    first ``spin1_right`` right nuclei are spin-1 (d = 9).  Generator (acts on row-major vec(rho)):
    H (x) 1 - 1 (x) H^T for Zeeman + hyperfine + exchange, and the Haberkorn sink -i k/2 (Q (x) 1 + 1 (x) Q^T);
    non-Hermitian -> Arnoldi, conserve_norm False.  Time unit: ns, energies in rad/ns."""
    rng = np.random.default_rng(seed)
    mults = [2] * n_left + [4] + [3 if i < spin1_right else 2 for i in range(n_right)]
    e = n_left
    n = len(mults)
    dims = [m * m for m in mults]
    gamma_e = 0.176  # rad / (ns mT)
    B0 = 0.5  # mT
    a_hf = rng.uniform(0.1, 1.0, n) * gamma_e  # hyperfine couplings, rad/ns
    J = 0.02
    kS, kT = 1.0e-3, 1.0e-3
    sx, sy, sz = _spin_ops(2)
    e2 = np.eye(2)
    S1 = [np.kron(o, e2) for o in (sx, sy, sz)]
    S2 = [np.kron(e2, o) for o in (sx, sy, sz)]
    S1S2 = sum(a @ b for a, b in zip(S1, S2, strict=True))
    Qs = 0.25 * np.eye(4) - S1S2
    Qt = np.eye(4) - Qs

    def left(op):
        return np.kron(op, np.eye(op.shape[0]))

    def right(op):
        return np.kron(np.eye(op.shape[0]), op.T)

    terms = []

    def add_comm(coef, ops: dict):
        terms.append((coef, {p: left(o) for p, o in ops.items()}))
        terms.append((-coef, {p: right(o) for p, o in ops.items()}))

    add_comm(gamma_e * B0, {e: S1[2] + S2[2]})
    add_comm(-2.0 * J, {e: S1S2})
    for p in range(n):
        if p == e:
            continue
        Ix, Iy, Iz = _spin_ops(mults[p])
        S = S1 if p < e else S2
        for c, I_c in enumerate((Ix, Iy, Iz)):
            add_comm(a_hf[p], {p: I_c, e: S[c]})
    for k, Q in ((kS, Qs), (kT, Qt)):
        terms.append((-0.5j * k, {e: left(Q)}))
        terms.append((-0.5j * k, {e: right(Q)}))
    cores = sop_to_mpo(dims, terms)
    hartree = []
    for p in range(n):
        rho = Qs if p == e else np.eye(mults[p])
        hartree.append(rho.reshape(-1).astype(complex))
    return Workload(f"c4_radical_pair_{n}site_D{D}", dims, {_key_full(n): cores}, hartree, D, dt * units.au_in_fs,
                    space="liouville", integrator="arnoldi", conserve_norm=False,
                    description=f"radical pair, {n_left}+{n_right} nuclear spins, Liouville-space MPDO (d = 4/9/16), "
                                f"w={max(c.shape[-1] for c in cores)}, D={D}, Arnoldi",
                    meta={"w": [c.shape[-1] for c in cores[:-1]], "electron_site": e})


def vibronic_chain(nsite: int = 128, N: int = 8, D: int = 512, dt_fs: float = 0.05, seed: int = 1238) -> Workload:
    """c5: ``nsite - 1`` HO-DVR modes (d = N) + one 2-level exciton site at the right end: diagonal potential MPO
    (w <= 4, last core 4-index) + kinetic MPO (w = 2) ending on the last vibrational site, i.e. the MPO pattern
    of tests/test_exiciton_propagate.py tiled along the chain."""
    rng = np.random.default_rng(seed)
    nv = nsite - 1
    freqs = rng.uniform(800.0, 3200.0, nv)
    prims = [HarmonicOscillator(N, f, units="cm-1") for f in freqs]
    dE, J = 0.01, 0.001
    kappa = rng.normal(0.0, 1.0e-4, nv)
    lam = rng.normal(0.0, 1.0e-4, nv)
    ex = Exciton(2)
    a = ex.get_annihilation_matrix()
    n1, n0, sx = a.T @ a, a @ a.T, a + a.T
    dims = [N] * nv + [2]
    terms = [(dE, {nv: n1}), (J, {nv: sx})]
    for i, p in enumerate(prims):
        q = np.array(p.get_grids())
        w = freqs[i] / units.au_in_cm1
        terms.append((1.0, {i: np.diag(0.5 * w**2 * q**2)}))
        terms.append((kappa[i], {i: np.diag(q), nv: n1}))
        terms.append((lam[i], {i: np.diag(q), nv: sx}))
    full = sop_to_mpo(dims, terms)
    pot = diag_cores(full[:-1]) + [full[-1]]
    kin = construct_kinetic_mpo(prims)
    key_pot = tuple((i,) for i in range(nv)) + ((nv, nv),)
    key_kin = tuple((i, i) for i in range(nv))
    hartree = [p.get_unitary()[0].reshape(1, N, 1).astype(complex) for p in prims] + [[0.0, 1.0]]
    return Workload(f"c5_vibronic_chain_{nsite}site_D{D}", dims, {key_pot: pot, key_kin: kin}, hartree, D, dt_fs,
                    description=f"{nv} HO-DVR modes (d={N}) + exciton site, diag potential MPO w<="
                                f"{max(c.shape[-1] for c in pot)} + kinetic MPO w=2, D={D}")


def h2co(path: str | None = None) -> Workload:
    """c1: the H2CO 6-mode grid MPO (diagonal potential cores from the reference's tests/h2co.tensor + kinetic MPO, HO-DVR
    d = 5, D = 16) exactly as committed in tests/golden/h2co_D16.npz (inputs made by tests/golden/make_golden.py)."""
    import ast
    import os

    if path is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "h2co_D16.npz")
    z = dict(np.load(path))
    n = len(z["dims"])
    operators = {}
    for ik in range(int(z["nkeys"])):
        key = ast.literal_eval(str(z[f"key{ik}"]))
        key = tuple(k if isinstance(k, tuple) else (k,) for k in key)
        cores, ic = [], 0
        while f"key{ik}_core{ic}" in z:
            cores.append(z[f"key{ik}_core{ic}"])
            ic += 1
        operators[key] = cores
    hartree = [z[f"init{i}"] for i in range(n)]   # the reference's own right-canonical initial MPS (3-D cores)
    return Workload("c1_h2co_6mode_D16", [int(d) for d in z["dims"]], operators, hartree, int(z["bond_dim"]),
                    float(z["dt_au"]) * units.au_in_fs, coupleJ=complex(z["coupleJ"]),
                    description="H2CO 6-mode grid-based DVR MPO (tests/h2co.tensor), HO-DVR d=5, D=16")


def by_name(name: str, **overrides) -> Workload:
    table = {"c1": h2co, "c2": henon_heiles, "c3": pyrazine_lvc, "c4": radical_pair, "c5": vibronic_chain}
    if name not in table:
        raise KeyError(f"unknown workload {name!r}; choose from {sorted(table)}")
    return table[name](**overrides)
