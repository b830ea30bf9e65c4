"""Checkpoint interoperability with the reference's ``wf_<jobname>.pkl`` files and MPO import from ``.npz``.

The reference saves a ``dill`` pickle of its live ``WFunc`` object (``Simulator.save_wavefunction``,
pytdscf/simulator_cls.py:577-589) and restarts by ``dill.load`` + ``WFunc(wf.ci_coef, wf.spf_coef, ints_prim)``
(:501-507).  This module reads such a file WITHOUT the reference package being installed (every ``pytdscf.*`` class is
resolved to a state-holding stand-in while unpickling) and writes files that the reference's restart path accepts
(objects whose pickled class references are ``pytdscf.wavefunction.WFunc``, ``pytdscf._mps_mpo.MPSCoefMPO``,
``pytdscf._mps_cls.LatticeInfo``, ``pytdscf._site_cls.SiteCoef`` and ``pytdscf._spf_cls.SPFCoef`` with exactly the
attributes those classes carry after ``alloc_random``; the cached environment blocks are left ``None`` and are rebuilt
by the reference at its first sweep, pytdscf/_mps_cls.py:836-845).

``import_mpo_npz`` / ``export_mpo_npz`` read and write MPO cores in the ``.npz`` layout of ``pympo.utils.import_npz`` /
``export_npz`` (pympo 0.1.8, used by the reference's docs/notebook/singlet_fission_nprocs.py:46): one array per core
under the keys ``W0 .. W{n-1}`` in site order.
"""
from __future__ import annotations

import pickle
import sys
import types

import numpy as np

_REF_TOP = ("pytdscf", "discvar", "jax", "jaxlib")
_NDARRAY_SUBCLASSES = {"myndarray"}   # pytdscf/_mps_mpo.py:1045-1061 (ndarray with an ``is_identity`` flag)


class _RefObject:
    """Stand-in for an instance of a reference class: unpickling fills ``__dict__`` with the saved state."""

    def __repr__(self):  # pragma: no cover
        return f"<{type(self).__module__}.{type(self).__name__} {sorted(self.__dict__)}>"


class _RefArray(np.ndarray):
    pass


def _stub_unpickler(file):
    try:
        import dill

        base = dill.Unpickler
    except ImportError:  # plain pickles of the same objects load without dill
        base = pickle.Unpickler
    cache: dict = {}

    class Unpickler(base):
        def find_class(self, module, name):
            if module.split(".")[0] in _REF_TOP:
                if name in _NDARRAY_SUBCLASSES:
                    return _RefArray
                key = (module, name)
                if key not in cache:
                    cache[key] = type(name, (_RefObject,), {"__module__": module})
                return cache[key]
            return super().find_class(module, name)

    return Unpickler(file)


def read_reference_wavefunction(path: str) -> dict:
    """Load a reference ``wf_*.pkl`` -> {"cores": [ndarray (D_l, d, D_r)], "gauges": [...], "nstate": int, "dims": [...]}.
    Only MPS-SM wavefunctions with one electronic "state" block are supported (the scope of ``backend="cuda"``)."""
    with open(path, "rb") as f:
        wf = _stub_unpickler(f).load()
    ci = getattr(wf, "ci_coef", None)
    if ci is None or not hasattr(ci, "superblock_states"):
        raise ValueError(f"{path} does not hold a reference WFunc with MPS coefficients")
    blocks = ci.superblock_states
    if len(blocks) != 1:
        raise NotImplementedError(f"{path} has {len(blocks)} state blocks; backend='cuda' supports nstate == 1")
    cores = [np.ascontiguousarray(np.asarray(s.data), dtype=np.complex128) for s in blocks[0]]
    gauges = [str(s.gauge) for s in blocks[0]]
    return {"cores": cores, "gauges": gauges, "nstate": 1, "dims": [int(c.shape[1]) for c in cores]}


def is_reference_pickle(path: str) -> bool:
    """True when the file's first pickled global is a ``pytdscf`` class (a reference ``WFunc`` dump)."""
    import pickletools

    with open(path, "rb") as f:
        head = f.read(4096)
    try:
        for op, arg, _pos in pickletools.genops(head):
            if op.name in ("GLOBAL", "STACK_GLOBAL", "SHORT_BINUNICODE", "BINUNICODE", "UNICODE") and isinstance(arg, str):
                if arg.startswith("pytdscf"):
                    return True
                if op.name == "GLOBAL":
                    return False
    except Exception:
        pass
    return b"pytdscf.wavefunction" in head


# ---------------------------------------------------------------------------------------------------------
# writing
# ---------------------------------------------------------------------------------------------------------
_WRITE_CLASSES = [("pytdscf.wavefunction", "WFunc"), ("pytdscf._mps_mpo", "MPSCoefMPO"), ("pytdscf._mps_cls", "LatticeInfo"),
                  ("pytdscf._site_cls", "SiteCoef"), ("pytdscf._spf_cls", "SPFCoef")]


def write_reference_wavefunction(path: str, cores: list, gauges: list[str]) -> str:
    """Write ``cores`` (site tensors, C-order complex128) and their gauge labels as a pickle the reference's restart path
    loads as a ``WFunc`` (pytdscf/simulator_cls.py:501-507).  The class references are written by name; the classes
    themselves are never imported here, so the file can be produced on a machine without the reference."""
    n = len(cores)
    if len(gauges) != n:
        raise ValueError("one gauge label per site tensor is required")
    saved = {}
    made = {}
    try:
        # temporary stand-in modules so that pickle can emit "module.Class" references (it verifies that the name resolves)
        for module, name in _WRITE_CLASSES:
            parts = module.split(".")
            for i in range(1, len(parts) + 1):
                mname = ".".join(parts[:i])
                if mname not in saved:
                    saved[mname] = sys.modules.get(mname)
                    sys.modules[mname] = types.ModuleType(mname)
            cls = type(name, (), {"__module__": module, "__qualname__": name})
            setattr(sys.modules[module], name, cls)
            made[name] = cls

        def obj(name, **state):
            o = made[name].__new__(made[name])
            o.__dict__.update(state)
            return o

        dims = [int(np.asarray(c).shape[1]) for c in cores]
        sites = [obj("SiteCoef", data=np.ascontiguousarray(np.asarray(c), dtype=np.complex128), gauge=str(g), isite=i)
                 for i, (c, g) in enumerate(zip(cores, gauges, strict=True))]
        lattice = obj("LatticeInfo", nspf_list_sites=[[d] for d in dims], nsite=n, dim_of_sites=list(dims), ndof_per_sites=[1] * n)
        ci = obj("MPSCoefMPO", op_sys_sites_dipo=None, ints_site_dipo=None, ints_site=None, op_sys_sites=None,
                 dofs_cas=list(range(n)), lattice_info_states=[lattice], superblock_states=[sites], nstate=1, nsite=n,
                 ndof_per_sites=[1] * n, site_is_dof=True, reshape_mat={})
        spf = obj("SPFCoef", data=[[np.eye(d, dtype=np.complex128) for d in dims]], nstate=1, ndof=n)
        wf = obj("WFunc", ci_coef=ci, spf_coef=spf, ints_prim=None)
        with open(path, "wb") as f:
            pickle.dump(wf, f, protocol=4)
    finally:
        for mname, old in saved.items():
            if old is None:
                sys.modules.pop(mname, None)
            else:
                sys.modules[mname] = old
    return path


# ---------------------------------------------------------------------------------------------------------
# MPO cores <-> npz  (pympo.utils.import_npz / export_npz layout)
# ---------------------------------------------------------------------------------------------------------
def import_mpo_npz(path: str) -> list[np.ndarray]:
    """MPO cores ``W0 .. W{n-1}`` of a ``.npz`` archive in site order: 4-index (w_l, d, d, w_r) or 3-index diagonal
    (w_l, d, w_r) arrays, ready for ``Model(operators={"hamiltonian": cores})`` / ``TensorOperator(mpo=cores)``."""
    z = np.load(path)
    keys = []
    for prefix in ("W", "arr_"):          # pympo's named layout, or np.savez's positional one
        keys = sorted((k for k in z.files if k.startswith(prefix) and k[len(prefix):].isdigit()), key=lambda k, p=prefix: int(k[len(p):]))
        if keys and [int(k[len(prefix):]) for k in keys] == list(range(len(keys))) and len(keys) == len(z.files):
            break
        keys = []
    if not keys:
        raise ValueError(f"{path}: expected arrays W0 .. W(n-1) (or arr_0 .. arr_(n-1)), found {sorted(z.files)}")
    cores = [np.asarray(z[k]) for k in keys]
    for i, c in enumerate(cores):
        if c.ndim not in (3, 4):
            raise ValueError(f"{path}: core {i} has {c.ndim} indices; MPO cores have 3 (diagonal) or 4")
        if i > 0 and cores[i - 1].shape[-1] != c.shape[0]:
            raise ValueError(f"{path}: MPO bond between cores {i - 1} and {i} does not match")
    if cores[0].shape[0] != 1 or cores[-1].shape[-1] != 1:
        raise ValueError(f"{path}: the outer MPO bonds must have dimension 1")
    return cores


def export_mpo_npz(path: str, cores: list) -> str:
    np.savez(path, **{f"W{i}": np.asarray(c) for i, c in enumerate(cores)})
    return path


__all__ = ["read_reference_wavefunction", "write_reference_wavefunction", "is_reference_pickle", "import_mpo_npz",
           "export_mpo_npz"]
