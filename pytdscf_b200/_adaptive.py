"""Rank-adaptive one-site TDVP (the reference's experimental ``adaptive=True`` mode, "A1TDVP").

Mirrors ``pytdscf/_mps_cls.py:862-1012`` (the adaptive branches of ``propagate_along_sweep``), ``:1921-2287``
(``get_psi_sigvec_psi_fullblock``, ``get_rank_and_projection_error``, ``get_op_block_full``,
``get_adaptive_rank_and_block``), ``:3660-3766`` (``truncate_op_block``, ``get_superblock_full``,
``get_actual_delta_rank``, ``is_max_rank``) and ``pytdscf/_site_cls.py:294-407`` (``thin_to_full``).

Before a site is propagated the next bond may grow by up to ``dD`` directions taken from the orthogonal complement of
the neighbouring isometry (full-QR completion); how many are kept is decided by the reference's projection-error
criterion f(D) = |H(D', D) Psi_left|^2 - |K(D) sigma|^2 + |H(D, D') Psi_right|^2.

Device strategy: every rectangular object of the reference (an environment block with an enlarged bra index, an
effective Hamiltonian whose output tensor is larger than its input -- ``SplitStack.stack(extend=True)`` /
``split(truncate=True)``) is a SQUARE operator on zero-padded operands, so the existing kernels (``tdvp_env_update``,
``tdvp_heff_apply``, ``tdvp_keff_apply``, ``tdvp_krylov_expm``, ``tdvp_qr_shift``) are used unchanged; the only
adaptive-specific library state is the Krylov size override (the reference sizes the Krylov space by the tensor BEFORE
extension, ``_integrator.py:178-186``).  ``thin_to_full`` is the Householder completion of the zero-padded isometry,
which reproduces LAPACK's full-QR columns (tests/test_gpu_kernels.py::test_qr_shift_zero_padded_matches_lapack_completion).
"""
from __future__ import annotations

import torch

from ._mps_cuda import Block, SiteCoef


def pad_to(t: torch.Tensor, shape: tuple[int, ...]) -> torch.Tensor:
    """Zero-pad ``t`` at the end of every axis up to ``shape`` (contiguous copy; identity when nothing grows)."""
    if tuple(t.shape) == tuple(shape):
        return t
    out = torch.zeros(shape, dtype=t.dtype, device=t.device)
    out[tuple(slice(0, n) for n in t.shape)] = t
    return out


# ---------------------------------------------------------------------------------------------------------
# full-rank neighbours
# ---------------------------------------------------------------------------------------------------------
def thin_to_full(eng, site: SiteCoef, delta_rank: int) -> SiteCoef:
    """A (l, c, r) -> (l, c, r + dr) or B (l, c, r) -> (l + dl, c, r): the isometry followed by ``delta_rank`` columns of
    the orthogonal complement in LAPACK's full-QR order (reference ``SiteCoef.thin_to_full``)."""
    l, c, r = site.data.shape  # noqa: E741
    if site.gauge == "A":
        dr = min(delta_rank, l * c - r)
        if dr <= 0:
            return SiteCoef(site.data.clone(), "A", site.isite)
        Q, _ = eng.qr_shift("A", pad_to(site.data, (l, c, r + dr)))          # rows (l, c) as in mat.reshape(l*c, r)
        Q = Q.clone()
        Q[:, :, :r] = site.data                                              # the reference re-aligns the signs of Q1
        return SiteCoef(Q.contiguous(), "A", site.isite)
    if site.gauge == "B":
        dl = min(delta_rank, c * r - l)
        if dl <= 0:
            return SiteCoef(site.data.clone(), "B", site.isite)
        # mat = data.reshape(l, c*r).T : rows ordered (c, r), one column per left-bond state
        X = pad_to(site.data, (l + dl, c, r)).permute(1, 2, 0).contiguous()   # (c, r, l + dl)
        Q, _ = eng.qr_shift("A", X)
        full = Q.permute(2, 0, 1).contiguous()                               # (l + dl, c, r)
        full[:l] = site.data
        return SiteCoef(full, "B", site.isite)
    raise ValueError(f"Invalid gauge: {site.gauge}")


def get_actual_delta_rank(sites: list[SiteCoef], isite: int, delta_rank: int) -> int:
    core = sites[isite]
    n = len(sites)
    if core.gauge == "A":
        if isite == n - 1:
            return 0
        l1, c1, r1 = core.shape
        l2, c2, r2 = sites[isite + 1].shape
        return max(min(delta_rank, min(l1 * c1 - r1, c2 * r2 - l2)), 0)
    if core.gauge == "B":
        if isite == 0:
            return 0
        l1, c1, r1 = core.shape
        l2, c2, r2 = sites[isite - 1].shape
        return max(min(delta_rank, min(c1 * r1 - l1, l2 * c2 - r2)), 0)
    raise ValueError(f"core.gauge={core.gauge} is not valid")


def get_superblock_full(eng, sites: list[SiteCoef], delta_rank: int) -> list[SiteCoef]:
    full = []
    for isite, core in enumerate(sites):
        if core.gauge == "Psi":
            full.append(SiteCoef(core.data.clone(), "Psi", core.isite))
        else:
            full.append(thin_to_full(eng, core, get_actual_delta_rank(sites, isite, delta_rank)))
    return full


def is_max_rank(site: SiteCoef, to: str, Dmax: int) -> bool:
    L, C, R = site.shape
    if to == "->":
        return L * C <= R or R >= Dmax
    return L >= C * R or L >= Dmax


# ---------------------------------------------------------------------------------------------------------
# rectangular environment blocks through the square kernels
# ---------------------------------------------------------------------------------------------------------
def _env_rect(eng, gauge: str, bra: torch.Tensor, ket: torch.Tensor, E, core, out=None):
    """contract_with_site_mpo with different bra / ket tensors: the ket is zero-padded to the bra's shape, the square
    kernel runs, and the ket index of the result is cut back.  Returns (D_bra, w, D_ket)."""
    D_ket = ket.shape[2] if gauge == "A" else ket.shape[0]
    res = eng.env_update(gauge, bra, pad_to(ket, tuple(bra.shape)), E, core)
    res = res[:, :, :D_ket].contiguous()
    if out is not None:
        out += res
        return out
    return res


def renormalize_braket(mps, psite: int, blocks: dict, H, A_is_sys: bool, bra: SiteCoef, ket: SiteCoef) -> dict:
    """``renormalize_op_psite`` with ``superblock_states_bra`` (reference _mps_mpo.py:421-696): same bookkeeping as
    ``MPSCoefCuda.renormalize_op_psite``; the overlap block is contracted explicitly and loses its identity flag."""
    eng = mps.eng
    gauge = "A" if A_is_sys else "B"
    b, k = bra.data, ket.data
    same = b is k or tuple(b.shape) == tuple(k.shape) and bool(torch.equal(b, k))
    if same:
        return mps.renormalize_op_psite(psite, blocks, H, A_is_sys, site=ket)
    nxt: dict = {}
    ov: Block = blocks["ovlp"]
    E_ovlp = None if ov.is_identity else ov.data
    nxt["ovlp"] = Block(_env_rect(eng, gauge, b, k, E_ovlp, None), False)
    for term in H.calc_point[psite + mps.site_offset]:
        if (term.is_left and A_is_sys) or (term.is_right and not A_is_sys):
            E = E_ovlp
        else:
            E = blocks[term.key]
        if (term.is_right and A_is_sys) or (term.is_left and not A_is_sys):
            if "summed" in nxt:
                _env_rect(eng, gauge, b, k, E, term.core, out=nxt["summed"])
            else:
                nxt["summed"] = _env_rect(eng, gauge, b, k, E, term.core)
        else:
            nxt[term.key] = _env_rect(eng, gauge, b, k, E, term.core)
    if "summed" in blocks:
        if "summed" in nxt:
            _env_rect(eng, gauge, b, k, blocks["summed"], None, out=nxt["summed"])
        else:
            nxt["summed"] = _env_rect(eng, gauge, b, k, blocks["summed"], None)
    return nxt


def truncate_op_block(blocks: dict, D: int, mode: str) -> dict:
    out = {}
    for key, val in blocks.items():
        if key == "ovlp":
            if val.is_identity:
                raise ValueError("an identity overlap block cannot be truncated (full-rank blocks are explicit)")
            t = val.data
        else:
            t = val
        if t.shape[0] < D or (mode == "braket" and t.shape[2] < D):
            raise ValueError(f"block of shape {tuple(t.shape)} is smaller than D={D} for key={key}")
        cut = t[:D].contiguous() if mode == "bra" else t[:D, :, :D].contiguous()
        out[key] = Block(cut, False) if key == "ovlp" else cut
    return out


def _square_terms_h(terms, Dl: int, Dr: int):
    """Pad the (possibly rectangular) L / R blocks of H_eff terms to (Dl, w, Dl) / (Dr, w, Dr)."""
    out = []
    for L, core, R, coef in terms:
        Lp = None if L is None else pad_to(L, (Dl, L.shape[1], Dl))
        Rp = None if R is None else pad_to(R, (Dr, R.shape[1], Dr))
        out.append((Lp, core, Rp, coef))
    return out


def _square_terms_k(terms, Dl: int, Dr: int):
    out = []
    for t in terms:
        L, R, coef = t[:3]
        Lp = None if L is None else pad_to(L, (Dl, L.shape[1], Dl))
        Rp = None if R is None else pad_to(R, (Dr, R.shape[1], Dr))
        out.append((Lp, Rp, coef))
    return out


def heff_apply_rect(eng, terms, psi: torch.Tensor, shape_out: tuple[int, int, int]) -> torch.Tensor:
    """multiplyH(...).dot with ``tensor_shapes_out``: (l, c, r) -> (L, c, R), L >= l, R >= r."""
    L, C, R = shape_out
    return eng.heff_apply(_square_terms_h(terms, L, R), pad_to(psi, (L, C, R)))


def keff_apply_rect(eng, terms, sigma: torch.Tensor, shape_out: tuple[int, int]) -> torch.Tensor:
    L, R = shape_out
    return eng.keff_apply(_square_terms_k(terms, L, R), pad_to(sigma, (L, R)))


# ---------------------------------------------------------------------------------------------------------
# the rank decision
# ---------------------------------------------------------------------------------------------------------
def get_rank_and_projection_error(mps, psite: int, Dmax: int, p: float, op_sys_full: dict, op_sys_thin: dict,
                                  op_env_full: dict, op_env_thin: dict, H, psi_left: torch.Tensor, sigvec: torch.Tensor,
                                  psi_right: torch.Tensor, to: str) -> tuple[int, float]:
    eng = mps.eng
    Dmin1, Dmin2 = sigvec.shape
    Dleft, d_left, _ = psi_left.shape
    _, d_right, Dright = psi_right.shape
    Dmax = min(Dmax, Dleft * d_left, Dright * d_right)
    Dmin = min(Dmin1, Dmin2)
    if Dmin > Dmax:
        raise ValueError(f"Dmin={Dmin} > Dmax={Dmax}")
    if Dmin == Dmax:
        return Dmin, 0.0
    fwd = to == "->"
    op_env_D = truncate_op_block(op_env_full, Dmax, "bra")
    op_sys_D = truncate_op_block(op_sys_full, Dmax, "bra")
    t1 = mps.operators_for_superH(psite if fwd else psite - 1, op_sys_thin if fwd else op_sys_D,
                                  op_env_D if fwd else op_env_thin, H, fwd)
    t2 = mps.operators_for_superH(psite + 1 if fwd else psite, op_sys_D if fwd else op_sys_thin,
                                  op_env_thin if fwd else op_env_D, H, fwd)
    tk = mps.operators_for_superK(op_sys_D, op_env_D, H, fwd)
    out_left = heff_apply_rect(eng, t1, psi_left, (Dleft, d_left, Dmax))
    out_right = heff_apply_rect(eng, t2, psi_right, (Dmax, d_right, Dright))
    out_sig = keff_apply_rect(eng, tk, sigvec, (Dmax, Dmax))
    # |x[..., :D]|^2 for every D at once (cumulative sums on the device, Dmax numbers to the host)
    nl = torch.cumsum((out_left.real**2 + out_left.imag**2).sum(dim=(0, 1)), 0).cpu().numpy()
    nr = torch.cumsum((out_right.real**2 + out_right.imag**2).sum(dim=(1, 2)), 0).cpu().numpy()
    a2 = out_sig.real**2 + out_sig.imag**2
    nk = torch.diagonal(torch.cumsum(torch.cumsum(a2, 0), 1)).cpu().numpy()
    total_prev = 0.0
    D = Dmin
    for D in range(Dmin, Dmax + 1):
        total = float(nl[D - 1]) - float(nk[D - 1]) + float(nr[D - 1])
        if D > Dmin:
            metric = (total - total_prev) / total
            if metric < p:
                return D - 1, metric
        total_prev = total
    return max(Dmin, D), 0.0


def get_adaptive_rank_and_block(mps, psite: int, full: list[SiteCoef], op_env_previous: dict, H, to: str, cfg):
    """Reference ``get_adaptive_rank_and_block``: returns (newD, error, op_env with newD bra states, op_env with newD bra
    and ket states) and enlarges the neighbouring isometry of ``mps.sites`` to newD."""
    eng = mps.eng
    sites = mps.sites
    fwd = to == "->"
    nb = psite + 1 if fwd else psite - 1
    # environment of the neighbour contracted with its full-rank version: bra and ket / bra only
    op_env_full_braket = renormalize_braket(mps, nb, op_env_previous, H, not fwd, full[nb], full[nb])
    op_env_full_bra = renormalize_braket(mps, nb, op_env_previous, H, not fwd, full[nb], sites[nb])
    def explicit(blocks, D):     # identity overlap (bra == ket, canonical) written out, so that it can be truncated
        if not blocks["ovlp"].is_identity:
            return blocks
        eye = torch.eye(D, dtype=torch.complex128, device=eng.torch_device).reshape(D, 1, D).contiguous()
        return dict(blocks, ovlp=Block(eye, False))

    Dnb = full[nb].shape[0] if fwd else full[nb].shape[2]
    op_env_full_braket = explicit(op_env_full_braket, Dnb)
    op_env_full_bra = explicit(op_env_full_bra, Dnb)
    op_sys_thin = mps.op_sys_sites[-1]
    if fwd:
        Dmax = min(cfg.Dmax, full[nb].shape[0])
        delta_rank = Dmax - sites[psite].shape[2]
    else:
        Dmax = min(cfg.Dmax, full[nb].shape[2])
        delta_rank = Dmax - sites[psite].shape[0]
    # Psi, sigma, Psi' and the system block of the full-rank isometry of this site
    psi = sites[psite]
    if fwd:
        A_data, sigvec = eng.qr_shift("A", psi.data)
        A_site = SiteCoef(A_data, "A", psite)
        psi_prime = eng.absorb("A", sigvec, sites[nb].data)
        A_full = thin_to_full(eng, A_site, delta_rank)
        op_sys_full_bra = renormalize_braket(mps, psite, op_sys_thin, H, True, A_full, A_site)
        psi_left, psi_right = psi.data, psi_prime
    else:
        B_data, sigvec = eng.qr_shift("B", psi.data)
        B_site = SiteCoef(B_data, "B", psite)
        psi_prime = eng.absorb("B", sigvec, sites[nb].data)
        B_full = thin_to_full(eng, B_site, delta_rank)
        op_sys_full_bra = renormalize_braket(mps, psite, op_sys_thin, H, False, B_full, B_site)
        psi_left, psi_right = psi_prime, psi.data
    op_sys_full_bra = explicit(op_sys_full_bra, A_site.shape[2] if fwd else B_site.shape[0])
    newD, error = get_rank_and_projection_error(mps, psite, Dmax, cfg.p_proj, op_sys_full_bra, op_sys_thin, op_env_full_bra,
                                                op_env_previous, H, psi_left, sigvec, psi_right, to)
    op_env_D_bra = truncate_op_block(op_env_full_bra, newD, "bra")
    op_env_D_braket = truncate_op_block(op_env_full_braket, newD, "braket")
    if fwd:
        sites[nb] = SiteCoef(full[nb].data[:newD].contiguous(), sites[nb].gauge, nb)
    else:
        sites[nb] = SiteCoef(full[nb].data[:, :, :newD].contiguous(), sites[nb].gauge, nb)
    return newD, error, op_env_D_bra, op_env_D_braket
