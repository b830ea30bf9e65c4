"""Run the UNMODIFIED reference (PyTDSCF 1.3.3, NumPy backend) on a ``pytdscf_b200.workloads.Workload`` and time it --
TEST / BENCH INFRASTRUCTURE (``bench.py --impl reference`` and the ``cpu_baseline`` leg; never imported by the product).

The reference package comes from ``oracle/reference_loader.py``: /root/reference in the build container, the
byte-compiled build product ``oracle/_ref`` on the GPU box.  Nothing of the reference is modified; two methods are
WRAPPED at run time to read a clock (``exp_superH_propagation_direct`` marks the start of every site update).

* ``time_full_steps``       the reference's own ``MPSCoefMPO.propagate`` (2 half sweeps) on the whole chain; first step
                            (environment bootstrap + opt_einsum expression caching) excluded -- SURVEY 8(d).
* ``time_bounded_sweep``    for workloads whose full step costs minutes of host time (config 4, D = 1024): the reference's own
                            ``propagate_along_sweep(begin_site=0, end_site=k)`` on the real state, real Krylov counts, timed
                            per site; the sites beyond k are filled in from the measured site of the same shape class
                            (a chain's second half mirrors the first), which ``extrapolation_check`` validates on a workload
                            where the full step is affordable.
"""
from __future__ import annotations

import os
import tempfile
import time

import numpy as np


def _load():
    from oracle.reference_loader import load_reference, reference_root

    root = reference_root(allow_built=True)
    load_reference(allow_built=True)
    return root


def source_kind() -> str | None:
    """'source tree' (build container) | 'oracle/_ref bytecode' (GPU box) | None (reference not available)."""
    from oracle.reference_loader import BUILT_ROOT, reference_root

    root = reference_root(allow_built=True)
    if root is None:
        return None
    return "oracle/_ref bytecode" if root == BUILT_ROOT else "source tree"


def build_reference_model(wl):
    """The reference's own ``Model`` for a workload: Exciton primitives of the site dimensions, one ``TensorOperator`` per
    MPO key (given as cores), Hartree-product initial weights."""
    _load()
    from pytdscf.basis import Exciton
    from pytdscf.dvr_operator_cls import TensorOperator
    from pytdscf.hamiltonian_cls import TensorHamiltonian
    from pytdscf.model_cls import Model

    basis = [Exciton(nstate=int(d)) for d in wl.dims]
    pot = {}
    for key, cores in wl.operators.items():
        rkey, legs = [], []
        for ind in key:
            if isinstance(ind, tuple) and len(ind) == 2:
                rkey.append((int(ind[0]), int(ind[1])))
                legs += [int(ind[0]), int(ind[1])]
            else:
                i = int(ind[0]) if isinstance(ind, tuple) else int(ind)
                rkey.append(i)
                legs.append(i)
        pot[tuple(rkey)] = TensorOperator(mpo=[np.asarray(c) for c in cores], legs=tuple(legs))
    if wl.coupleJ != 0:
        pot[()] = complex(wl.coupleJ)
    ham = TensorHamiltonian(ndof=len(basis), potential=[[pot]], kinetic=None, backend="numpy")
    model = Model(basis, {"hamiltonian": ham}, bond_dim=int(wl.bond_dim), space=wl.space)
    model.init_HartreeProduct = [[np.asarray(h) for h in wl.hartree]]
    return model


class _Session:
    """const.set_runtype + MPSCoefMPO.alloc_random exactly as Simulator.propagate does it (simulator_cls.py:247-283,
    :495-545), without the property/IO layers."""

    def __init__(self, wl, thresh_sil: float = 1e-9):
        self.wl = wl
        self.model = build_reference_model(wl)
        from pytdscf import _helper
        from pytdscf._const_cls import const
        from pytdscf._mps_mpo import MPSCoefMPO

        self._cwd = os.getcwd()
        self._tmp = tempfile.TemporaryDirectory()
        os.chdir(self._tmp.name)
        const.set_runtype(jobname="refrun", space=wl.space, integrator=wl.integrator, conserve_norm=wl.conserve_norm,
                          dvr=self.model.basinfo.is_DVR, verbose=0, thresh_sil=thresh_sil)
        _helper._Debug.niter_krylov.clear()
        _helper._Debug.site_now = 0
        self.helper = _helper
        self.mps = MPSCoefMPO.alloc_random(self.model)
        self.matH = self.model.hamiltonian
        self.site_marks: list[tuple[int, float]] = []
        self._orig = MPSCoefMPO.exp_superH_propagation_direct
        marks = self.site_marks
        orig = self._orig

        def marked(this, *a, **k):
            marks.append((int(_helper._Debug.site_now), time.perf_counter()))
            return orig(this, *a, **k)

        MPSCoefMPO.exp_superH_propagation_direct = marked
        self._cls = MPSCoefMPO

    def close(self):
        self._cls.exp_superH_propagation_direct = self._orig
        os.chdir(self._cwd)
        self._tmp.cleanup()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def krylov_counts(self) -> dict:
        return dict(self.helper._Debug.niter_krylov)


def time_full_steps(wl, nsteps: int, warm_steps: int = 1) -> dict:
    """Seconds per full time step of the reference's MPSCoefMPO.propagate (mean over ``nsteps`` after ``warm_steps``)."""
    with _Session(wl) as s:
        times = []
        for i in range(warm_steps + nsteps):
            t0 = time.perf_counter()
            s.mps.propagate(wl.dt_au, None, s.matH)
            times.append(time.perf_counter() - t0)
        timed = times[warm_steps:]
        return {"seconds_per_step": float(np.mean(timed)), "step_seconds": [float(t) for t in times],
                "sweeps_per_s": 2.0 / float(np.mean(timed)), "krylov": s.krylov_counts()}


def _shape_class(wl, p: int):
    """(min bond, physical dimension, max bond) of site p: the mirror-symmetric cost class of its site update."""
    from pytdscf_b200._mps_cuda import bond_dims

    dl, dr = bond_dims(wl.dims, p, wl.bond_dim)
    return (min(dl, dr), int(wl.dims[p]), max(dl, dr))


def time_bounded_sweep(wl, end_site: int, warm_full_steps: int = 0) -> dict:
    """The reference's own forward half sweep ``propagate_along_sweep(begin_site=0, end_site=end_site)`` on the real state
    (right environments of the whole chain prebuilt by its ``construct_op_sites`` and handed over through ``op_sys_sites``,
    as between two sweeps of a real run; that construction is the one-off bootstrap and is not timed).  Returns the seconds
    of every site update 0 .. end_site-1 and the estimate of a whole sweep: measured sites + every remaining site charged
    with the mean measured time of its shape class (sites without a measured class: scaled by algorithmic flops from the
    nearest class of the same bonds)."""
    n = len(wl.dims)
    with _Session(wl) as s:
        mps, matH = s.mps, s.matH
        for _ in range(warm_full_steps):
            mps.propagate(wl.dt_au, None, matH)
        if mps.op_sys_sites is None:
            mps.ints_site = mps.get_ints_site(None)
            mps.matH_sweep = mps.get_matH_sweep(matH)
            t0 = time.perf_counter()
            mps.op_sys_sites = mps.construct_op_sites(mps.superblock_states, ints_site=mps.ints_site, begin_site=n - 1,
                                                      end_site=0, matH_cas=mps.matH_sweep)
            t_env = time.perf_counter() - t0
        else:
            t_env = 0.0
        # hand over exactly the end_site + 1 right environments this bounded sweep consumes (the reference checks the count)
        mps.op_sys_sites = mps.op_sys_sites[-(end_site + 1):]
        s.site_marks.clear()
        t0 = time.perf_counter()
        mps.propagate_along_sweep(mps.ints_site, mps.matH_sweep, wl.dt_au, begin_site=0, end_site=end_site)
        t_end = time.perf_counter()
        marks = s.site_marks
        per_site = {}
        for i, (p, t) in enumerate(marks):
            nxt = marks[i + 1][1] if i + 1 < len(marks) else t_end
            per_site[p] = nxt - t
        # the end site of a bounded sweep only gets its H solve: not a full site update, drop it from the classes
        measured = {p: t for p, t in per_site.items() if p != end_site}
        classes: dict = {}
        for p, t in measured.items():
            classes.setdefault(_shape_class(wl, p), []).append(t)
        cmean = {c: float(np.mean(v)) for c, v in classes.items()}

        def flops(c):
            lo, d, hi = c
            return d * (lo * lo * hi + lo * hi * hi)

        total, filled = 0.0, []
        for p in range(n):
            if p in measured:
                total += measured[p]
                continue
            c = _shape_class(wl, p)
            if c in cmean:
                est = cmean[c]
            else:
                same = [k for k in cmean if (k[0], k[2]) == (c[0], c[2])] or list(cmean)
                k = min(same, key=lambda k: abs(np.log(flops(k) / flops(c))))
                est = cmean[k] * flops(c) / flops(k)
            filled.append((p, c, est))
            total += est
        return {"sweep_seconds_estimate": total, "sweeps_per_s": 1.0 / total, "measured_seconds": float(t_end - t0),
                "env_bootstrap_seconds": t_env, "site_seconds": {int(p): float(t) for p, t in sorted(measured.items())},
                "filled_sites": [(int(p), [int(x) for x in c], float(e)) for p, c, e in filled],
                "krylov": s.krylov_counts(), "end_site": int(end_site)}
