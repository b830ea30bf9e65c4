"""The switch points of INTEGRATION.md as executable code: route the narrow waist of an imported, UNMODIFIED PyTDSCF through
an engine of this package (``pytdscf_b200._engine.Engine`` on a GPU; the tests inject the oracle's NumPy engine).

    import pytdscf                                   # the reference
    from pytdscf_b200.reference_adapter import install
    undo = install(pytdscf, engine, level="integrators")
    ... pytdscf.Simulator(...).propagate(...)        # its own host logic, this package's kernels
    undo()

What is replaced (reference file:line), nstate = 1, TensorHamiltonian (MPO) models:
  level "contractions"
    multiplyH_MPS_direct_MPO._op_lcr_dot   _contraction.py:1038-1178   -> engine.heff_apply (one term)
    multiplyK_MPS_direct_MPO._op_lr_dot    _contraction.py:1297-1356   -> engine.keff_apply (one term)
    contract_with_site_mpo                 _contraction.py:148-397     -> engine.env_update
    SiteCoef.gauge_trf (no regularisation) _site_cls.py:138-292        -> engine.qr_shift
  level "integrators" (adds)
    short_iterative_lanczos / _arnoldi     _integrator.py:453-655, 287-432 -> engine.krylov_expm with the terms of the
                                           multiplyOp object; the reference's warm-up table ``_Debug.niter_krylov`` is read
                                           and written exactly where the reference does (_integrator.py:178-186, 642-652)
The reference keeps every array as a host ndarray, so this adapter copies operands to the engine and results back around
each call: it demonstrates that the waist is sufficient and numerically equivalent, it is not the fast path (that is
``pytdscf_b200.Simulator``, which keeps the whole state on the device).  Nothing else in this package imports this module.
"""
from __future__ import annotations

import numpy as np


def _to_np(t):
    return t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)


class _Waist:
    def __init__(self, ref, eng):
        self.ref, self.eng = ref, eng
        self._cores: dict = {}

    # -- operands -----------------------------------------------------------------------------------------
    def block(self, b):
        """Environment block: int (the reference's identity placeholder) -> None, ndarray (2 or 3 indices) -> device."""
        if isinstance(b, (int, np.integer)):
            return None
        a = np.asarray(b)
        if a.ndim == 2:                       # overlap-type block without an MPO bond: a bond of dimension 1 for the engine
            a = a.reshape(a.shape[0], 1, a.shape[1])
        return self.eng.to_device(a)

    def core(self, c):
        """Site operator: int -> None; OperatorCore (3-index diagonal / 4-index) or a bare d x d matrix -> device core."""
        if isinstance(c, (int, np.integer)):
            return None
        data = getattr(c, "data", c)
        if isinstance(data, (int, np.integer)):
            return None
        key = id(data)
        hit = self._cores.get(key)
        if hit is None or hit[0] is not data:          # uploaded once per core object, like OperatorCore.apply_backend would
            a = np.asarray(data)
            if a.ndim == 2:
                a = a.reshape(1, a.shape[0], a.shape[1], 1)
            hit = (data, self.eng.upload_core(a))
            self._cores[key] = hit
        return hit[1]

    # -- contractions ---------------------------------------------------------------------------------------
    def op_lcr_dot(self, _self, op_l, op_c, op_r, trial, key=None):
        out = self.eng.heff_apply([(self.block(op_l), self.core(op_c), self.block(op_r), 1.0)], self.eng.to_device(trial))
        return _to_np(out)

    def op_lr_dot(self, _self, op_l, op_r, trial, key=None):
        if isinstance(op_l, (int, np.integer)) and isinstance(op_r, (int, np.integer)):
            return trial
        out = self.eng.keff_apply([(self.block(op_l), self.block(op_r), 1.0)], self.eng.to_device(trial))
        return _to_np(out)

    def contract_with_site_mpo(self, mat_bra, mat_ket, op_LorR, op_site):
        gauge = mat_bra.gauge
        if gauge not in ("A", "B"):
            raise ValueError(f"contract_with_site_mpo: gauge {gauge!r}")
        out = self.eng.env_update(gauge, self.eng.to_device(np.asarray(mat_bra.data)), self.eng.to_device(np.asarray(mat_ket.data)),
                                  self.block(op_LorR), self.core(op_site))
        return _to_np(out)

    def gauge_trf(self, original):
        waist = self
        SiteCoef = self.ref._site_cls.SiteCoef

        def gauge_trf(site, key, regularize=False):
            if regularize or key not in ("Psi2Asigma", "C2Asigma", "Psi2sigmaB", "C2sigmaB"):
                return original(site, key, regularize)
            g = "A" if key in ("Psi2Asigma", "C2Asigma") else "B"
            new, sigma = waist.eng.qr_shift(g, waist.eng.to_device(np.asarray(site.data)))
            return SiteCoef(data=_to_np(new), gauge=g, isite=site.isite), _to_np(sigma)

        return gauge_trf

    # -- integrators ----------------------------------------------------------------------------------------
    def terms_of(self, multiplyOp):
        """(hterms, kterms) of a reference multiplyH / multiplyK object for state block (0, 0), coupleJ as the overlap term."""
        ham = multiplyOp.matH_cas
        coupleJ = complex(ham.coupleJ[0][0])
        if hasattr(multiplyOp, "op_lcr_states"):
            ops = multiplyOp.op_lcr_states[0][0]
            terms = []
            for key, (l, c, r) in ops.items():
                if key == "ovlp":
                    if coupleJ != 0.0:
                        terms.append((self.block(l), None, self.block(r), coupleJ))
                else:
                    terms.append((self.block(l), self.core(c), self.block(r), 1.0))
            return terms, None
        ops = multiplyOp.op_lr_states[0][0]
        terms = []
        for key, (l, r) in ops.items():
            if key == "ovlp":
                if coupleJ != 0.0:
                    terms.append((self.block(l), self.block(r), coupleJ))
            else:
                terms.append((self.block(l), self.block(r), 1.0))
        return None, terms

    def krylov(self, kind):
        waist = self
        integ = self.ref._integrator
        const = self.ref._const_cls.const
        dbg = integ._Debug

        def solve(scale, multiplyOp, psi_states, thresh):
            if len(psi_states) != 1:
                raise NotImplementedError("reference_adapter: nstate == 1 only")
            _, _, n_warmup = integ._iter_info(psi_states)
            hterms, kterms = waist.terms_of(multiplyOp)
            psi = waist.eng.to_device(np.asarray(psi_states[0]))
            n = waist.eng.krylov_expm(kind, complex(scale), float(thresh), int(n_warmup), bool(const.conserve_norm), psi,
                                      hterms=hterms, kterms=kterms)
            dbg.niter_krylov[dbg.site_now] = int(n)
            return [_to_np(psi)]

        return solve


def install(ref, engine, level: str = "integrators"):
    """Patch the imported reference package ``ref`` (module object of ``pytdscf``); returns a function that undoes it."""
    if level not in ("contractions", "integrators"):
        raise ValueError("level must be 'contractions' or 'integrators'")
    import importlib

    for sub in ("_contraction", "_integrator", "_site_cls", "_const_cls", "_mps_mpo", "_mps_cls"):
        importlib.import_module(f"{ref.__name__}.{sub}")
    w = _Waist(ref, engine)
    con, integ, site = ref._contraction, ref._integrator, ref._site_cls
    saved = []

    def patch(obj, name, new):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, new)

    patch(con.multiplyH_MPS_direct_MPO, "_op_lcr_dot", lambda self, l, c, r, t, key=None: w.op_lcr_dot(self, l, c, r, t, key))
    patch(con.multiplyK_MPS_direct_MPO, "_op_lr_dot", lambda self, l, r, t, key=None: w.op_lr_dot(self, l, r, t, key))
    original_env = con.contract_with_site_mpo
    for mod in (con, ref._mps_mpo, ref._mps_cls):         # the name is imported into the modules that call it
        if getattr(mod, "contract_with_site_mpo", None) is original_env:
            patch(mod, "contract_with_site_mpo", w.contract_with_site_mpo)
    patch(site.SiteCoef, "gauge_trf", w.gauge_trf(site.SiteCoef.gauge_trf))
    if level == "integrators":
        patch(integ, "short_iterative_lanczos", w.krylov("lanczos"))
        patch(integ, "short_iterative_arnoldi", w.krylov("arnoldi"))

    def undo():
        while saved:
            obj, name, old = saved.pop()
            setattr(obj, name, old)

    return undo


__all__ = ["install"]
