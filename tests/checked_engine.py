"""Debug / test helper: run every kernel call on the CUDA ``Engine`` AND on the oracle's NumPy kernels with the same
inputs, record the largest deviation per operation (gauge-invariant where the result is not unique), and hand the GPU
result on.  Pinpoints which device kernel departs from the reference inside a full sweep.  Test infrastructure only."""
from __future__ import annotations

import numpy as np
import torch

from oracle.oracle_engine import OracleEngine
from pytdscf_b200._engine import DeviceCore


def _c(t):
    return None if t is None else t.detach().cpu()


def _core(c):
    if c is None:
        return None
    return DeviceCore(_c(c.data), None)


class CheckedEngine:
    def __init__(self, eng):
        self.eng = eng
        self.orc = OracleEngine(strict=False)      # replays calls the CUDA engine has already accepted
        self.torch_device = eng.torch_device
        self.dev: dict[str, float] = {}
        self.worst: dict[str, tuple] = {}

    def _rec(self, name, err, info=None):
        err = float(err)
        if err > self.dev.get(name, -1.0):
            self.dev[name] = err
            self.worst[name] = info

    def __getattr__(self, name):  # stats, reset_stats, to_device, upload_core, empty, gemm_profile ...
        return getattr(self.eng, name)

    def heff_apply(self, terms, psi):
        out = self.eng.heff_apply(terms, psi)
        ref = self.orc.heff_apply([(_c(L), _core(c), _c(R), k) for L, c, R, k in terms], _c(psi))
        self._rec("heff_apply", (out.cpu() - ref).abs().max() / max(1e-300, ref.abs().max()), tuple(psi.shape))
        return out

    def keff_apply(self, terms, sigma):
        out = self.eng.keff_apply(terms, sigma)
        ref = self.orc.keff_apply([(_c(t[0]), _c(t[1]), t[2]) for t in terms], _c(sigma))
        self._rec("keff_apply", (out.cpu() - ref).abs().max() / max(1e-300, ref.abs().max()), tuple(sigma.shape))
        return out

    def env_update(self, gauge, bra, ket, E, core, out=None, accumulate=False):
        before = _c(out).clone() if (out is not None and accumulate) else None
        res = self.eng.env_update(gauge, bra, ket, E, core, out=out, accumulate=accumulate)
        ref = self.orc.env_update(gauge, _c(bra), _c(ket), _c(E), _core(core))
        if before is not None:
            ref = ref + before
        self._rec("env_update", (res.cpu() - ref).abs().max() / max(1e-300, ref.abs().max()), (gauge, tuple(ket.shape)))
        return res

    def krylov_expm(self, kind, scale, thresh, n_warmup, conserve_norm, psi, *, hterms=None, kterms=None, size_override=None):
        x0 = _c(psi).clone()
        n = self.eng.krylov_expm(kind, scale, thresh, n_warmup, conserve_norm, psi, hterms=hterms, kterms=kterms,
                                 size_override=size_override)
        kw = {"size_override": size_override}
        if hterms is not None:
            kw["hterms"] = [(_c(L), _core(c), _c(R), k) for L, c, R, k in hterms]
        else:
            kw["kterms"] = [(_c(t[0]), _c(t[1]), t[2]) for t in kterms]
        n_ref = self.orc.krylov_expm(kind, scale, thresh, n_warmup, conserve_norm, x0, **kw)
        self._rec("krylov_expm", (psi.cpu() - x0).abs().max() / max(1e-300, x0.abs().max()), (tuple(psi.shape), n, n_ref))
        self._rec("krylov_niter", abs(n - n_ref), (tuple(psi.shape), n, n_ref))
        return n

    def lanczos_eigvec(self, psi, hterms, root=0, thresh=1e-9):
        return self.eng.lanczos_eigvec(psi, hterms, root, thresh)

    def qr_shift(self, gauge, psi, regularize=False):
        site, sigma = self.eng.qr_shift(gauge, psi, regularize=regularize)
        rs, rsig = self.orc.qr_shift(gauge, _c(psi), regularize=regularize)
        s, g = site.cpu().numpy(), sigma.cpu().numpy()
        if gauge == "A":
            prod, rprod = np.tensordot(s, g, axes=(2, 0)), np.tensordot(rs.numpy(), rsig.numpy(), axes=(2, 0))
            m = s.reshape(-1, s.shape[2])
            iso = np.abs(m.conj().T @ m - np.eye(m.shape[1])).max()
        else:
            prod, rprod = np.tensordot(g, s, axes=(1, 0)), np.tensordot(rsig.numpy(), rs.numpy(), axes=(1, 0))
            m = s.reshape(s.shape[0], -1)
            iso = np.abs(m @ m.conj().T - np.eye(m.shape[0])).max()
        tag = "qr_shift_reg" if regularize else "qr_shift"
        import os
        if iso > 1e-12 and os.environ.get("CHECKED_DUMP"):
            np.savez(os.path.join(os.environ["CHECKED_DUMP"], f"qr_iso_{gauge}_{'x'.join(map(str, psi.shape))}_{iso:.1e}.npz"),
                     psi=psi.cpu().numpy(), site=s, sigma=g, ref_site=rs.numpy(), ref_sigma=rsig.numpy())
        self._rec(tag + "_product", np.abs(prod - rprod).max() / max(1e-300, np.abs(rprod).max()), (gauge, tuple(psi.shape)))
        self._rec(tag + "_isometry", iso, (gauge, tuple(psi.shape)))
        self._rec(tag + "_direct", np.abs(s - rs.numpy()).max(), (gauge, tuple(psi.shape)))
        return site, sigma

    def svd(self, M):
        U, s, Vh = self.eng.svd(M)
        Ur, sr, Vr = self.orc.svd(_c(M))
        u, v = U.cpu().numpy(), Vh.cpu().numpy()
        self._rec("svd_values", np.abs(s - sr).max(), tuple(M.shape))
        self._rec("svd_reconstruct", np.abs((u * s[None, :]) @ v - M.cpu().numpy()).max(), tuple(M.shape))
        self._rec("svd_U_orth", np.abs(u.conj().T @ u - np.eye(u.shape[1])).max(), tuple(M.shape))
        self._rec("svd_V_orth", np.abs(v @ v.conj().T - np.eye(v.shape[0])).max(), tuple(M.shape))
        return U, s, Vh

    def svd_truncate(self, sigma, p, keepdim=False, regularize=False):
        U, S, Vh, r = self.eng.svd_truncate(sigma, p, keepdim=keepdim, regularize=regularize)
        Ur, Sr, Vr, rr = self.orc.svd_truncate(_c(sigma), p, keepdim=keepdim, regularize=regularize)
        prod = U.cpu().numpy() @ S.cpu().numpy() @ Vh.cpu().numpy()
        rprod = Ur.numpy() @ Sr.numpy() @ Vr.numpy()
        self._rec("svd_truncate_product", np.abs(prod - rprod).max(), (tuple(sigma.shape), r, rr))
        self._rec("svd_truncate_S", np.abs(S.cpu().numpy() - Sr.numpy()).max(), (tuple(sigma.shape), r, rr))
        self._rec("svd_truncate_rank", abs(r - rr), (tuple(sigma.shape), r, rr))
        u = U.cpu().numpy()
        self._rec("svd_truncate_U_orth", np.abs(u.conj().T @ u - np.eye(u.shape[1])).max(), tuple(sigma.shape))
        return U, S, Vh, r

    def pinv(self, X, rcond=1e-13):
        out = self.eng.pinv(X, rcond)
        ref = self.orc.pinv(_c(X), rcond)
        self._rec("pinv", (out.cpu() - ref).abs().max() / max(1e-300, ref.abs().max()), (tuple(X.shape), rcond))
        return out

    def zgemm(self, A, B, transA=0, transB=0, alpha=1.0, beta=0.0, C_out=None):
        before = _c(C_out).clone() if C_out is not None else None
        out = self.eng.zgemm(A, B, transA, transB, alpha, beta, C_out)
        ref = self.orc.zgemm(_c(A), _c(B), transA, transB, alpha, beta, before)
        self._rec("zgemm", (out.cpu() - ref).abs().max() / max(1e-300, ref.abs().max()), (tuple(A.shape), tuple(B.shape)))
        return out

    def absorb(self, gauge, sigma, site):
        out = self.eng.absorb(gauge, sigma, site)
        ref = self.orc.absorb(gauge, _c(sigma), _c(site))
        self._rec("absorb", (out.cpu() - ref).abs().max() / max(1e-300, ref.abs().max()), (gauge, tuple(sigma.shape), tuple(site.shape)))
        return out

    def inner(self, bra, ket, conj=True):
        out = self.eng.inner(bra, ket, conj)
        ref = self.orc.inner(_c(bra), _c(ket), conj)
        self._rec("inner", abs(out - ref) / max(1e-300, abs(ref)), (tuple(bra.shape), conj))
        return out

    def overlap_site(self, bra, ket, block, conj_bra):
        out = self.eng.overlap_site(bra, ket, block, conj_bra)
        ref = self.orc.overlap_site(_c(bra), _c(ket), _c(block), conj_bra)
        self._rec("overlap_site", (out.cpu() - ref).abs().max() / max(1e-300, ref.abs().max()), (tuple(bra.shape), conj_bra))
        return out

    def report(self) -> str:
        return "\n".join(f"  {k:24s} {v:.3e}   worst: {self.worst[k]}" for k, v in sorted(self.dev.items()))
