"""``DeviceNumericsEngine``: the oracle's NumPy engine with the SVD-based calls replaced by a NumPy restatement of what
``pytdscf_b200/csrc/svd.cu`` does on the device -- one-sided Jacobi with zgesvj's threshold sqrt(m) eps, columns below 1e-20 of the
largest one frozen, a sweep without a rotation above 8 x threshold ends the iteration (60 sweeps at most, otherwise an error),
left vectors of numerically-zero singular values taken from an orthonormal completion, ``tdvp_svd_truncate`` / ``tdvp_pinv`` /
``Engine.regularize_site`` built on it exactly as the library and ``_engine.py`` build them.

Test infrastructure (CPU container, no GPU): it shows whether HOST logic that has only ever run on LAPACK's conventions also
holds with the device's -- different singular-vector phases, different vectors inside (near-)null spaces, completion vectors
instead of noise vectors.  It is NOT the device kernel (pair order and summation order differ), so it pins nothing; the GPU
parity tests do that."""
from __future__ import annotations

import numpy as np

from oracle import tdvp_oracle as orc
from oracle.oracle_engine import OracleEngine, _chk

NEGLIGIBLE, SIGNIFICANT, EPS = 1.0e-20, 8.0, 2.220446049250313e-16
SQRT_EPSRHO = 1.0e-4


def jacobi_svd(A: np.ndarray):
    """Thin SVD of an m x n matrix (m >= n) the way svd_exec does it; returns (U, s, Vh, sweeps)."""
    m, n = A.shape
    assert m >= n
    G = A.astype(complex).copy()
    V = np.eye(n, dtype=complex)
    tol = np.sqrt(m) * EPS
    frozen2 = NEGLIGIBLE * NEGLIGIBLE * (np.abs(G) ** 2).sum(axis=0).max()
    converged, last_cos, sweeps = False, 0.0, 0
    for sweep in range(60):
        nrot, maxcos = 0, 0.0
        for p in range(n - 1):
            for q in range(p + 1, n):
                x, y = G[:, p].copy(), G[:, q].copy()
                al, be = np.vdot(x, x).real, np.vdot(y, y).real
                g = np.vdot(x, y)
                ga = abs(g)
                lim = tol * np.sqrt(al) * np.sqrt(be)
                if not (al > frozen2 and be > frozen2 and ga > lim):
                    continue
                if ga > SIGNIFICANT * lim:
                    nrot += 1
                    maxcos = max(maxcos, ga / (np.sqrt(al) * np.sqrt(be)))
                ph = g / ga
                zeta = (be - al) / (2.0 * ga)
                tt = (1.0 if zeta >= 0 else -1.0) / (abs(zeta) + np.sqrt(1.0 + zeta * zeta))
                c = 1.0 / np.sqrt(1.0 + tt * tt)
                s = c * tt
                yt = y * np.conj(ph)
                G[:, p], G[:, q] = c * x - s * yt, s * x + c * yt
                vx, vy = V[:, p].copy(), V[:, q] * np.conj(ph)
                V[:, p], V[:, q] = c * vx - s * vy, s * vx + c * vy
        sweeps = sweep + 1
        if nrot == 0:
            converged = True
            break
        last_cos = maxcos
    if not converged and not (0.0 < last_cos < 1.0e-10):
        raise RuntimeError(f"svd: one-sided Jacobi did not converge in 60 sweeps (largest |cos| {last_cos:.3e})")
    norms = np.sqrt((np.abs(G) ** 2).sum(axis=0))
    order = np.argsort(-norms, kind="stable")
    s = norms[order]
    r = 0
    for i in range(n):
        if s[i] > 0.0 and s[i] > 4.0 * NEGLIGIBLE * s[0]:
            r = i + 1
    U = np.zeros((m, n), dtype=complex)
    for i in range(r):
        U[:, i] = G[:, order[i]] / s[i]
    if r < n:                                  # complete_columns: Householder QR of [U_r | 0]
        full = np.linalg.qr(U[:, :r], mode="complete")[0] if r > 0 else np.eye(m, dtype=complex)
        U[:, r:] = full[:, r:n]
    return U, s, V[:, order].conj().T, sweeps


class DeviceNumericsEngine(OracleEngine):
    reorder_mpo_channels = True     # as the CUDA engine: DeviceMPO moves identity-prefix / -suffix channels first / last

    def __init__(self):
        super().__init__()
        self.svd_log: list = []     # (shape, sweeps)

    def _svd(self, A: np.ndarray):
        m, n = A.shape
        if m < n:                                                     # Engine.svd: wide matrices through the adjoint
            U2, s, Vh2, sw = jacobi_svd(np.ascontiguousarray(A.conj().T))
            U, Vh = Vh2.conj().T, U2.conj().T
        else:
            U, s, Vh, sw = jacobi_svd(A)
        self.svd_log.append((A.shape, sw))
        return U, s, Vh

    def svd(self, M):
        _chk(M, "M", 2)
        U, s, Vh = self._svd(M.numpy())
        return self._wrap(np.ascontiguousarray(U)), s, self._wrap(np.ascontiguousarray(Vh))

    def svd_truncate(self, sigma, p, keepdim=False, regularize=False):      # tdvp_svd_truncate + Engine.svd_truncate
        _chk(sigma, "sigma", 2)
        A = sigma.numpy()
        n = A.shape[0]
        if A.shape[0] != A.shape[1]:
            raise ValueError("svd_truncate: the bond matrix must be square")
        U, s, Vh = self._svd(A)
        cs = np.cumsum(s)
        idx = n
        for i in range(n):
            if cs[i] / cs[n - 1] >= 1.0 - p:
                idx = i + 1
                break
        thin = s[:idx].copy()
        if regularize and n != 1:
            thin = np.where(thin > SQRT_EPSRHO, thin, thin + SQRT_EPSRHO * np.exp(-thin / SQRT_EPSRHO))
        thin = thin / np.linalg.norm(thin)
        k = n if keepdim else idx
        S = np.zeros((k, k), dtype=complex)
        S[np.arange(idx), np.arange(idx)] = thin
        if keepdim:
            return self._wrap(U), self._wrap(S), self._wrap(Vh), idx
        return self._wrap(np.ascontiguousarray(U[:, :idx])), self._wrap(S), self._wrap(np.ascontiguousarray(Vh[:idx])), idx

    def pinv(self, X, rcond=1e-13):                                           # tdvp_pinv
        _chk(X, "X", 2)
        A = X.numpy()
        if A.shape[0] != A.shape[1]:
            raise ValueError("pinv: square matrices only")
        off = A - np.diag(np.diag(A))
        if not off.any():                                                     # diagonal probe: no SVD
            d = np.abs(np.diag(A))
            cut = rcond * d.max()
            inv = np.array([1.0 / v if abs(v) > cut else 0.0 for v in np.diag(A)])
            return self._wrap(np.diag(inv).astype(complex))
        U, s, Vh = self._svd(A)
        inv = np.where(s > rcond * s[0], 1.0 / np.where(s > 0, s, 1.0), 0.0)
        return self._wrap(np.ascontiguousarray((Vh.conj().T * inv[None, :]) @ U.conj().T))

    def regularize_site(self, psi):                                           # Engine.regularize_site
        A = psi.numpy()
        Dl, d, Dr = A.shape
        U, s, Vh = self._svd(np.ascontiguousarray(A.transpose(0, 2, 1).reshape(Dl * Dr, d)))
        s_reg = np.where(s > SQRT_EPSRHO, s, s + SQRT_EPSRHO * np.exp(-s / SQRT_EPSRHO))
        return np.ascontiguousarray(((U * s_reg[None, :]) @ Vh).reshape(Dl, Dr, d).transpose(0, 2, 1))

    def qr_shift(self, gauge, psi, regularize=False):
        _chk(psi, "psi", 3)
        Dl, d, Dr = psi.shape
        if (gauge == "A" and Dl * d < Dr) or (gauge != "A" and Dr * d < Dl):
            raise ValueError("QR / LQ shift needs a tall matricisation")
        A = self.regularize_site(psi) if regularize else psi.numpy()
        if gauge == "A":
            Q, s = orc.shift_qr(A, False)
            return self._wrap(Q), self._wrap(s)
        s, B = orc.shift_lq(A, False)
        return self._wrap(B), self._wrap(s)
