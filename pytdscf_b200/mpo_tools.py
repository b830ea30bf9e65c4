"""Input preparation: build full-length MPO core lists from sums of operator products.

The reference delegates this to the external ``pympo`` package (absent here); the hot
path only consumes the resulting ``list[np.ndarray]`` cores ``W[p]`` of shape (w_l, d_bra, d_ket, w_r)
(``pytdscf/model_cls.py:215-284``).  This is host-side, one-off setup and is not part of the GPU path.
"""
from __future__ import annotations

import numpy as np


def sop_to_mpo(dims: list[int], terms: list[tuple[complex, dict[int, np.ndarray]]]) -> list[np.ndarray]:
    """Finite-state-machine MPO of ``sum_t coef_t * prod_{p in t} O_t[p]``.

    Bond channels: "not started", "done", and one channel per product term that crosses the bond.
    Terms sharing the same left factor at the same starting site share a channel until they differ.
    """
    n = len(dims)
    cleaned = []
    for coef, ops in terms:
        if not ops:
            raise ValueError("scalar terms are not supported; add them to one site operator")
        sites = sorted(ops)
        if sites[0] < 0 or sites[-1] >= n:
            raise ValueError("operator site out of range")
        for s in sites:
            if ops[s].shape != (dims[s], dims[s]):
                raise ValueError(f"operator on site {s} has shape {ops[s].shape}, expected {(dims[s],) * 2}")
        cleaned.append((complex(coef), {s: np.asarray(ops[s], dtype=np.complex128) for s in sites}, sites[0], sites[-1]))

    # channel lists per bond b (between site b and b+1); index into `cleaned`
    crossing = [[t for t, (_, _, lo, hi) in enumerate(cleaned) if lo <= b < hi] for b in range(n - 1)]

    def chan(b: int) -> dict:
        if b < 0:
            return {"start": 0}
        if b >= n - 1:
            return {"done": 0}
        m = {"start": 0, "done": 1}
        for k, t in enumerate(crossing[b]):
            m[t] = 2 + k
        return m

    cores = []
    for p in range(n):
        left, right = chan(p - 1), chan(p)
        W = np.zeros((len(left), dims[p], dims[p], len(right)), dtype=np.complex128)
        eye = np.eye(dims[p])
        if "start" in left and "start" in right:
            W[left["start"], :, :, right["start"]] = eye
        if "done" in left and "done" in right:
            W[left["done"], :, :, right["done"]] = eye
        for t, (coef, ops, lo, hi) in enumerate(cleaned):
            if p < lo or p > hi:
                continue
            O = ops.get(p, eye)
            if lo == hi:
                W[left["start"], :, :, right["done"]] += coef * O
            elif p == lo:
                W[left["start"], :, :, right[t]] += coef * O
            elif p == hi:
                W[left[t], :, :, right["done"]] += O
            else:
                W[left[t], :, :, right[t]] += O
        cores.append(W)
    return cores


def mpo_to_dense(cores: list[np.ndarray]) -> np.ndarray:
    """Dense matrix of a full-length MPO (tests only; exponential in the chain length)."""
    acc = None
    for W in cores:
        if W.ndim == 3:
            d = W.shape[1]
            full = np.zeros((W.shape[0], d, d, W.shape[2]), dtype=np.complex128)
            idx = np.arange(d)
            full[:, idx, idx, :] = W
            W = full
        if acc is None:
            acc = W[0]  # (i, j, t)
        else:
            acc = np.einsum("ijc,ckls->ikjls", acc, W).reshape(acc.shape[0] * W.shape[1], acc.shape[1] * W.shape[2], W.shape[3])
    return acc[:, :, 0]
