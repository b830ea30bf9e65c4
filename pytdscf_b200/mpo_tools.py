"""Input preparation: build full-length MPO core lists from sums of operator products.

The reference delegates this to the external ``pympo`` package (absent here); the hot
path only consumes the resulting ``list[np.ndarray]`` cores ``W[p]`` of shape (w_l, d_bra, d_ket, w_r)
(``pytdscf/model_cls.py:215-284``).  This is host-side, one-off setup and is not part of the GPU path.
"""
from __future__ import annotations

import numpy as np


def _op_id(site: int, op: np.ndarray) -> tuple:
    return (site, op.shape, np.ascontiguousarray(op).tobytes())


def sop_to_mpo(dims: list[int], terms: list[tuple[complex, dict[int, np.ndarray]]]) -> list[np.ndarray]:
    """Finite-state-machine MPO of ``sum_t coef_t * prod_{p in t} O_t[p]``.

    Bond channels: "not started", "done", plus shared channels for the terms crossing a bond.  Two-operator
    terms (site a < site b) share a channel either with every term that has the same LEFT factor (a, O_a)
    ("prefix" channel; the coefficient is applied when the term ends) or with every term that has the same
    RIGHT factor (b, O_b) ("suffix" channel; the coefficient is applied when the term starts) -- whichever
    factor is shared by more terms.  Star-shaped couplings (many sites to one hub) therefore cost one channel
    per hub operator instead of one per term.  Terms with 3+ operators get a private channel.
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

    # channel identity of each multi-site term
    pre_count: dict = {}
    suf_count: dict = {}
    for coef, ops, lo, hi in cleaned:
        if len(ops) == 2:
            pre_count[_op_id(lo, ops[lo])] = pre_count.get(_op_id(lo, ops[lo]), 0) + 1
            suf_count[_op_id(hi, ops[hi])] = suf_count.get(_op_id(hi, ops[hi]), 0) + 1
    chan_of = []  # per term: (channel id, mode)
    for t, (coef, ops, lo, hi) in enumerate(cleaned):
        if lo == hi:
            chan_of.append((None, "single"))
        elif len(ops) == 2:
            pid, sid = _op_id(lo, ops[lo]), _op_id(hi, ops[hi])
            if suf_count[sid] >= pre_count[pid]:
                chan_of.append((("suf",) + sid, "suffix"))
            else:
                chan_of.append((("pre",) + pid, "prefix"))
        else:
            chan_of.append((("own", t), "own"))

    def chan(b: int) -> dict:
        if b < 0:
            return {"start": 0}
        if b >= n - 1:
            return {"done": 0}
        m = {"start": 0, "done": 1}
        for t, (coef, ops, lo, hi) in enumerate(cleaned):
            cid = chan_of[t][0]
            if cid is not None and lo <= b < hi and cid not in m:
                m[cid] = len(m)
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
        passed = set()
        opened = set()
        closed = set()
        for t, (coef, ops, lo, hi) in enumerate(cleaned):
            if p < lo or p > hi:
                continue
            cid, mode = chan_of[t]
            O = ops.get(p, eye)
            if mode == "single":
                W[left["start"], :, :, right["done"]] += coef * O
            elif p == lo:
                if mode == "suffix":
                    W[left["start"], :, :, right[cid]] += coef * O
                elif (cid, "open") not in opened:  # prefix / own: the shared left factor is placed once
                    W[left["start"], :, :, right[cid]] += O
                    opened.add((cid, "open"))
            elif p == hi:
                if mode == "prefix":
                    W[left[cid], :, :, right["done"]] += coef * O
                elif mode == "own":
                    W[left[cid], :, :, right["done"]] += coef * O
                elif (cid, "close") not in closed:  # suffix: the shared right factor is placed once
                    W[left[cid], :, :, right["done"]] += O
                    closed.add((cid, "close"))
            else:
                if (cid, p) not in passed:
                    W[left[cid], :, :, right[cid]] += O
                    passed.add((cid, p))
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


def sop_to_dense(dims: list[int], terms) -> np.ndarray:
    """Dense matrix of the operator sum (tests only)."""
    total = None
    for coef, ops in terms:
        mat = np.array([[1.0 + 0j]])
        for p, d in enumerate(dims):
            mat = np.kron(mat, np.asarray(ops.get(p, np.eye(d)), dtype=np.complex128))
        total = coef * mat if total is None else total + coef * mat
    return total
