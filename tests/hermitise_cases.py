"""Shared checker for the MPDO hermitisation tests (CPU host-logic test and GPU parity test): the before / after site
tensors of tests/golden/hermitise.npz (unmodified reference, tests/golden/make_golden_hermitise.py)."""
from __future__ import annotations

import os

import numpy as np

from tests.golden_io import GOLDEN_DIR


def load_cases():
    return dict(np.load(os.path.join(GOLDEN_DIR, "hermitise.npz")))


def dense(cores):
    """The full tensor of an MPS / MPDO chain (gauge invariant), small chains only."""
    t = np.asarray(cores[0])
    for c in cores[1:]:
        t = np.tensordot(t, np.asarray(c), axes=(-1, 0))
    return t.reshape(t.shape[1:-1])


def run_and_check(tag: str, eng, tol: float):
    import torch

    from pytdscf_b200._mps_cuda import MPSCoefCuda

    z = load_cases()
    n = 3
    before = [z[f"{tag}_before{i}"] for i in range(n)]
    after = [z[f"{tag}_after{i}"] for i in range(n)]
    mps = MPSCoefCuda(eng, [eng.to_device(np.ascontiguousarray(c)) for c in before])
    mps.op_sys_sites = ["stale"]
    mps.hermitise()
    got = [s.data.cpu().numpy() if isinstance(s.data, torch.Tensor) else np.asarray(s.data) for s in mps.sites]
    assert [g.shape for g in got] == [a.shape for a in after]
    assert [s.gauge for s in mps.sites] == [str(g) for g in z[f"{tag}_gauges"]]
    assert mps.op_sys_sites is None                       # cached environments dropped (_mps_cls.py:2310-2311)
    ref = dense(after)
    err = np.abs(dense(got) - ref).max() / np.abs(ref).max()
    assert err < tol, (tag, err)
    # right-canonical sites behind the centre
    for g in got[1:]:
        m = g.reshape(g.shape[0], -1)
        assert np.abs(m @ m.conj().T - np.eye(m.shape[0])).max() < 1e-12
    # the result is Hermitian to the accuracy of the truncation: exactly (rounding) where nothing was cut
    rho = dense(got)
    q = [int(round(np.sqrt(d))) for d in rho.shape]
    r = rho.reshape([x for d in q for x in (d, d)])
    perm = [p for i in range(len(q)) for p in (2 * i + 1, 2 * i)]
    asym = np.abs(r - r.conj().transpose(perm)).max() / np.abs(r).max()
    return err, asym
