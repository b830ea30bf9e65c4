"""CPU stand-in for ``pytdscf_b200._engine.Engine`` built on the oracle's NumPy kernels -- TEST INFRASTRUCTURE.

It exists so that the host-side logic of the product (``pytdscf_b200/_mps_cuda.py``: environment bookkeeping,
sweep order, Krylov warm-up history, Simulator loop, multi-process plumbing) can be exercised by ``-m "not gpu"``
tests and by world_size-2 ``gloo`` tests in a container without a GPU.  It is never imported by the package:
``Simulator`` always constructs the CUDA ``Engine`` and raises without a GPU.  Tensors are torch CPU complex128
tensors (zero-copy NumPy views)."""
from __future__ import annotations

import numpy as np
import torch

from oracle import tdvp_oracle as orc
from pytdscf_b200._engine import DeviceCore


def _np(t):
    return None if t is None else t.numpy()


def _core(core: DeviceCore | None):
    if core is None or core.data is None:
        return None
    return orc.SiteCore((), 0, core.data.numpy(), False, False)


def _chk(t, name: str, ndim: int | None = None):
    """The preconditions of the real engine (``pytdscf_b200/_engine.py::_chk_tensor``) and of the C ABI behind it, which reads raw
    memory: a contiguous complex128 torch tensor whose lazy conjugate bit is not set.  The NumPy kernels below would accept
    anything; enforcing the device's contract here makes every CPU host-logic test also a test of what the host hands to the ABI."""
    if t is None:
        return
    if not isinstance(t, torch.Tensor) or t.dtype != torch.complex128:
        raise TypeError(f"{name} must be a complex128 torch tensor (got {type(t).__name__}, {getattr(t, 'dtype', None)})")
    if not t.is_contiguous() or t.is_conj():
        raise TypeError(f"{name} must be contiguous with its conjugate bit resolved (the C ABI reads raw memory)")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name} must have {ndim} indices, got shape {tuple(t.shape)}")


def _unit_channel(block, ch: int):
    """Copy of an environment block (D, w, D) with channel ``ch`` replaced by the unit matrix -- what the device library uses for
    a channel named in ``id_channels`` (it copies instead of multiplying)."""
    b = block.clone()
    b[:, ch, :] = torch.eye(b.shape[0], dtype=b.dtype)
    return b


def _chk_shortcut(full: np.ndarray, short: np.ndarray, what: str):
    """The identity-channel promise of ``tdvp_heff_term.id_channels`` / ``tdvp_keff_term.id_channels``, stated as what it has to
    guarantee: the product with the promised channels taken as unit matrices equals the full contraction.  (Comparing the
    blocks themselves would be too strict: a zero-padded or sub-space-projected site leaves zeros on the diagonal of <B|B> for
    bond states that carry no amplitude, which changes nothing.)"""
    scale = max(float(np.abs(full).max()), 1.0e-300)
    dev = float(np.abs(full - short).max()) / scale
    if dev > 1.0e-9:
        raise ValueError(f"{what}: the identity-channel shortcut of the device library would change the result by {dev:.2e} (relative)")


def _chk_hterms(terms, psi):
    Dl, d, Dr = psi.shape
    for L, core, R, _coef in terms:
        _chk(L, "L", 3)
        _chk(R, "R", 3)
        if L is not None and (L.shape[0] != Dl or L.shape[2] != Dl):
            raise ValueError(f"left block {tuple(L.shape)} does not match psi {tuple(psi.shape)}")
        if R is not None and (R.shape[0] != Dr or R.shape[2] != Dr):
            raise ValueError(f"right block {tuple(R.shape)} does not match psi {tuple(psi.shape)}")
        if core is not None and core.data is not None:
            _chk(core.data, "W")
            if L is not None and L.shape[1] != core.wl or R is not None and R.shape[1] != core.wr:
                raise ValueError("block / core bond dimension mismatch")
            if core.data.shape[1] != d:
                raise ValueError("core / site physical dimension mismatch")


def _chk_kterms(terms, sigma):
    Dl, Dr = sigma.shape
    for L, R, *_rest in terms:
        _chk(L, "L", 3)
        _chk(R, "R", 3)
        if L is not None and (L.shape[0] != Dl or L.shape[2] != Dl) or R is not None and (R.shape[0] != Dr or R.shape[2] != Dr):
            raise ValueError("K_eff block does not match the bond matrix")
        if L is not None and R is not None and L.shape[1] != R.shape[1]:
            raise ValueError("K_eff term: MPO bond dimensions of L and R differ")


class OracleEngine:
    def __init__(self, strict: bool = True):
        """``strict``: enforce the device engine's preconditions and verify the identity-channel promise (CPU host-logic tests).
        ``tests/checked_engine.py`` replays calls the CUDA engine has ALREADY accepted and passes ``strict=False``: there the
        checks would add nothing, and the replay must never be what makes a GPU test fail."""
        self.strict = bool(strict)
        self.torch_device = torch.device("cpu")
        self._stats = {"solves": 0, "matvecs": 0, "flops": 0.0, "launches": 0}

    def close(self):
        pass

    def to_device(self, a) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128).copy())

    def upload_core(self, data) -> DeviceCore:
        if data is None:
            return DeviceCore(None)
        return DeviceCore(self.to_device(data), None)

    def empty(self, *shape) -> torch.Tensor:
        return torch.empty(shape, dtype=torch.complex128)

    def _wrap(self, a: np.ndarray) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(a))

    # -- precondition checks (strict mode only) ------------------------------------------------------------------
    def _pre(self, *specs):
        """specs: (tensor, name[, ndim]) triples checked against the device engine's contract."""
        if self.strict:
            for spec in specs:
                _chk(*spec)

    def _require(self, ok: bool, msg: str):
        if self.strict and not ok:
            raise ValueError(msg)

    def heff_apply(self, terms, psi):
        if self.strict:
            _chk(psi, "psi", 3)
            _chk_hterms(terms, psi)
        out = self._heff_sum(terms, psi)
        if self.strict:          # the identity-channel promise: the product must not change when the named channels are unit matrices
            promised = [(L, c, R, coef) for (L, c, R, coef) in terms
                        if c is not None and c.data is not None and ((L is not None and c.l_id >= 0) or (R is not None and c.r_id >= 0))]
            if promised:
                short = [(None if L is None else (_unit_channel(L, c.l_id) if c.l_id >= 0 else L), c,
                          None if R is None else (_unit_channel(R, c.r_id) if c.r_id >= 0 else R), coef) for (L, c, R, coef) in promised]
                _chk_shortcut(self._heff_sum(promised, psi), self._heff_sum(short, psi), "H_eff")
        return self._wrap(out)

    @staticmethod
    def _heff_sum(terms, psi) -> np.ndarray:
        out = None
        for L, c, R, coef in terms:
            add = orc.heff_term(_np(L), _core(c), _np(R), psi.numpy())
            if complex(coef) != 1.0:
                add = add * complex(coef)
            out = add.copy() if out is None else out + add
        return out

    def keff_apply(self, terms, sigma):
        if self.strict:
            _chk(sigma, "sigma", 2)
            _chk_kterms(terms, sigma)
        out = self._keff_sum(terms, sigma)
        if self.strict:
            promised = [t for t in terms if len(t) > 3 and t[3] is not None and t[0] is not None and t[1] is not None]
            if promised:
                short = [(_unit_channel(L, ids[0]) if ids[0] >= 0 else L, _unit_channel(R, ids[1]) if ids[1] >= 0 else R, coef)
                         for (L, R, coef, ids) in promised]
                _chk_shortcut(self._keff_sum(promised, sigma), self._keff_sum(short, sigma), "K_eff")
        return self._wrap(out)

    @staticmethod
    def _keff_sum(terms, sigma) -> np.ndarray:
        out = None
        for (L, R, coef, *_ids) in terms:
            add = orc.keff_term(_np(L), _np(R), sigma.numpy())
            if complex(coef) != 1.0:
                add = add * complex(coef)
            out = add.copy() if out is None else out + add
        return out

    def env_update(self, gauge, bra, ket, E, core, out=None, accumulate=False):
        self._pre((bra, "bra", 3), (ket, "ket", 3), (E, "E", 3), (out, "out", 3),
                  (None if core is None else core.data, "W"))
        self._require(tuple(bra.shape) == tuple(ket.shape),
                      "tdvp_env_update takes bra and ket of one shape (rectangular blocks are zero-padded by the caller)")
        res = orc.env_update_term(gauge, bra.numpy(), ket.numpy(), _np(E), _core(core))
        if out is not None and accumulate:
            out += self._wrap(res)
            return out
        return self._wrap(res)

    def krylov_expm(self, kind, scale, thresh, n_warmup, conserve_norm, psi, *, hterms=None, kterms=None, size_override=None):
        if self.strict:
            _chk(psi, "psi")
            if hterms is not None:
                _chk_hterms(hterms, psi)
            else:
                _chk_kterms(kterms, psi)
        last = n_warmup + 2 if n_warmup > 0 else 0
        if hterms is not None:
            mv = lambda x: self.heff_apply(hterms, self._wrap(x)).numpy()  # noqa: E731
        else:
            mv = lambda x: self.keff_apply(kterms, self._wrap(x)).numpy()  # noqa: E731
        solver = orc.sia_reference if kind == "arnoldi" else orc.sil_reference
        y, n = solver(complex(scale), mv, psi.numpy().copy(), thresh, last_niter=last, conserve_norm=conserve_norm,
                      maxsize=size_override)
        psi.copy_(self._wrap(y))
        self._stats["solves"] += 1
        self._stats["matvecs"] += n
        return n

    def lanczos_eigvec(self, psi, hterms, root=0, thresh=1e-9):
        if self.strict:
            _chk(psi, "psi", 3)
            _chk_hterms(hterms, psi)
        mv = lambda x: self.heff_apply(hterms, self._wrap(x)).numpy()  # noqa: E731
        y, n = orc.lanczos_ground_state(mv, psi.numpy().copy(), root, thresh)
        psi.copy_(self._wrap(y / np.linalg.norm(y)))
        return n

    def qr_shift(self, gauge, psi, regularize=False):
        self._pre((psi, "psi", 3))
        Dl, d, Dr = psi.shape
        # Engine.qr_shift / TDVP_ERR_SHAPE: tall matricisations only (the reference's economic QR would shrink the bond)
        self._require(Dl * d >= Dr if gauge == "A" else Dr * d >= Dl, "QR / LQ shift needs a tall matricisation")
        if gauge == "A":
            A, s = orc.shift_qr(psi.numpy(), regularize)
            return self._wrap(A), self._wrap(s)
        s, B = orc.shift_lq(psi.numpy(), regularize)
        return self._wrap(B), self._wrap(s)

    def svd(self, M):
        self._pre((M, "M", 2))
        U, s, Vh = np.linalg.svd(M.numpy(), full_matrices=False)
        return self._wrap(U), s, self._wrap(Vh)

    def svd_truncate(self, sigma, p, keepdim=False, regularize=False):
        self._pre((sigma, "sigma", 2))
        self._require(sigma.shape[0] == sigma.shape[1], "svd_truncate: the bond matrix must be square")        # tdvp_svd_truncate
        U, S, Vh, rank = orc.truncate_bond(None, sigma.numpy(), None, p, regularize, keepdim)
        return self._wrap(U), self._wrap(S.astype(complex)), self._wrap(Vh), rank

    def pinv(self, X, rcond=1e-13):
        self._pre((X, "X", 2))
        self._require(X.shape[0] == X.shape[1], "pinv: square matrices only")                                # tdvp_pinv
        return self._wrap(np.linalg.pinv(X.numpy(), rcond=rcond))

    def zgemm(self, A, B, transA=0, transB=0, alpha=1.0, beta=0.0, C_out=None):
        # Engine.zgemm passes shape[1] as the leading dimension: row-major contiguous operands only
        self._pre((A, "A", 2), (B, "B", 2), (C_out, "C", 2))
        a = {0: A.numpy(), 1: A.numpy().T, 2: A.numpy().conj().T}[transA]
        b = {0: B.numpy(), 1: B.numpy().T, 2: B.numpy().conj().T}[transB]
        out = alpha * (a @ b)
        if C_out is not None:
            C_out.copy_(self._wrap(out + beta * C_out.numpy()))
            return C_out
        return self._wrap(out)

    def absorb(self, gauge, sigma, site):
        self._pre((sigma, "sigma", 2), (site, "site", 3))
        self._require(sigma.shape[1] == site.shape[0] if gauge == "A" else sigma.shape[0] == site.shape[2],
                      "absorb: bond dimensions of the matrix and the site differ")
        if gauge == "A":
            return self._wrap(np.tensordot(sigma.numpy(), site.numpy(), axes=(1, 0)))
        return self._wrap(np.tensordot(site.numpy(), sigma.numpy(), axes=(2, 0)))

    def inner(self, bra, ket, conj=True):
        self._pre((bra, "bra"), (ket, "ket"))
        self._require(bra.numel() == ket.numel(), "inner: sizes differ")
        a = bra.numpy().ravel()
        return complex(np.inner(np.conj(a) if conj else a, ket.numpy().ravel()))

    def overlap_site(self, bra, ket, block, conj_bra):
        self._pre((bra, "bra", 3), (ket, "ket", 3), (block, "block", 2))
        b = np.conj(bra.numpy()) if conj_bra else bra.numpy()
        return self._wrap(np.einsum("abc,abk->ck", b, np.einsum("ibk,ai->abk", ket.numpy(), block.numpy())))

    def stats(self):
        return dict(self._stats)

    def reset_stats(self):
        self._stats = {"solves": 0, "matvecs": 0, "flops": 0.0, "launches": 0}
