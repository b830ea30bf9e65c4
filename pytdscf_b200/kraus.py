"""Host-side helpers for the Kraus-map workflow of open-system runs (reference: pytdscf/kraus.py:17-124, :434-470).

``lindblad_to_kraus`` turns a set of Lindblad jump operators {L_j} into the Kraus operators {B_q} of one time step of
the dissipator, exp(D dt) = sum_q B_q (x) conj(B_q) with D = sum_j [L_j (x) conj(L_j) - 1/2 (L_j^+ L_j (x) 1 + 1 (x) L_j^T
conj(L_j))] (row-major vectorisation of the density matrix; Phys. Rev. Lett. 116, 237201).  The result, shape (k, d, d),
is what ``Model(kraus_op={(site,): B})`` takes; the map itself is applied on the GPU (``MPSCoefCuda.apply_kraus``).
These are d^2 x d^2 matrices prepared once before a run -- setup code, NumPy/SciPy on the host, not on the hot path."""
from __future__ import annotations

from math import isqrt

import numpy as np
import scipy.linalg


def dissipator_superoperator(lindblad_ops: list) -> np.ndarray:
    """D acting on the row-major vectorised density matrix, vec(rho)[(a, b)] = rho[a, b]."""
    ops = [np.asarray(L, dtype=np.complex128) for L in lindblad_ops]
    if not ops or any(L.ndim != 2 or L.shape[0] != L.shape[1] for L in ops) or len({L.shape for L in ops}) != 1:
        raise ValueError("Lindblad operators must be square matrices of one common dimension")
    d = ops[0].shape[0]
    eye = np.eye(d)
    D = np.zeros((d * d, d * d), dtype=np.complex128)
    for L in ops:
        LdL = L.conj().T @ L
        D += np.kron(L, L.conj()) - 0.5 * (np.kron(LdL, eye) + np.kron(eye, LdL.T))
    return D


def supergate_to_kraus(G: np.ndarray, tol: float = 1e-14) -> np.ndarray:
    """Kraus operators of a completely positive map given as its matrix G on row-major vec(rho), G = sum_q B_q (x) conj(B_q):
    eigen-decomposition of the (Hermitian, positive semi-definite) Choi matrix C[(a, c), (b, e)] = G[(a, b), (c, e)];
    eigenvalues <= tol are dropped.  Returns an array (k, d, d)."""
    G = np.asarray(G)
    d = isqrt(G.shape[0])
    if d * d != G.shape[0] or G.shape[0] != G.shape[1]:
        raise ValueError("a superoperator on d x d density matrices is a d^2 x d^2 matrix")
    choi = G.reshape(d, d, d, d).transpose(0, 2, 1, 3).reshape(d * d, d * d)      # G[(a,b),(c,e)] = sum_q B[a,c] conj(B[b,e])
    choi = 0.5 * (choi + choi.conj().T)
    w, V = np.linalg.eigh(choi)
    keep = [i for i in range(len(w)) if w[i] > tol]
    if not keep:
        raise ValueError("the map has no positive Choi eigenvalue")
    return np.stack([np.sqrt(w[i]) * V[:, i].reshape(d, d) for i in keep], axis=0).astype(np.complex128)


def lindblad_to_kraus(lindblad_ops: list, dt: float, tol: float = 1e-14) -> np.ndarray:
    """Kraus operators B (k, d, d) of exp(D dt) for the dissipator D of ``lindblad_ops``; checks positivity and that the
    set reproduces the map to 1e-13 (as the reference asserts)."""
    if not dt > 0:
        raise ValueError("dt must be positive")
    G = scipy.linalg.expm(dissipator_superoperator(list(lindblad_ops)) * dt)
    B = supergate_to_kraus(G, tol)
    rebuilt = sum(np.kron(b, b.conj()) for b in B)
    if np.abs(rebuilt - G).max() > 1e-13 * max(1.0, np.abs(G).max()):
        raise ValueError("exp(D dt) is not completely positive to working precision (Choi matrix has negative eigenvalues)")
    return B


def trace_kraus_dim(rdm: np.ndarray, d: int) -> np.ndarray:
    """Trace the ancilla (Kraus) part out of a reduced density matrix whose index is (system d) x (ancilla K):
    (dK, dK) -> (d, d), or a time series (t, dK, dK) -> (t, d, d)."""
    rdm = np.asarray(rdm)
    dK = rdm.shape[-1]
    if dK % d:
        raise ValueError(f"Kraus dimension reduction: dK={dK} must be divisible by d={d}")
    K = dK // d
    if rdm.ndim == 2:
        return np.einsum("aKbK->ab", rdm.reshape(d, K, d, K))
    if rdm.ndim == 3:
        return np.einsum("taKbK->tab", rdm.reshape(-1, d, K, d, K))
    raise ValueError(f"rdm.ndim={rdm.ndim} must be 2 or 3")


__all__ = ["dissipator_superoperator", "supergate_to_kraus", "lindblad_to_kraus", "trace_kraus_dim"]
