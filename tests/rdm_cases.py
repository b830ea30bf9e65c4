"""Shared loader for the reduced-density goldens (tests/golden/rdm.npz, made by tests/golden/make_golden_rdm.py)."""
import os

import numpy as np

from tests.golden_io import GOLDEN_DIR


def rdm_cases():
    z = np.load(os.path.join(GOLDEN_DIR, "rdm.npz"))
    out = []
    for name in sorted({k.split("__")[0] for k in z.files}):
        k = 0
        while f"{name}__{k}__key" in z.files:
            out.append((name, tuple(int(v) for v in z[f"{name}__{k}__key"]), z[f"{name}__{k}__rdm"]))
            k += 1
    return out


def check_rdms(eng, load_run, MPSCoefCuda, atol):
    seen = {}
    for name, key, ref in rdm_cases():
        if name not in seen:
            g = load_run(name)
            seen[name] = MPSCoefCuda(eng, [eng.to_device(c) for c in g["final"]])
        got = seen[name].get_reduced_densities(key, space="liouville" if name.startswith("liouville") else "hilbert")[0]
        assert got.shape == ref.shape, (name, key, got.shape, ref.shape)
        np.testing.assert_allclose(got, ref, rtol=0, atol=atol, err_msg=f"{name} {key}")
    # list form = one array per key
    name, key, ref = [c for c in rdm_cases() if not c[0].startswith("liouville")][0]
    many = seen[name].get_reduced_densities([key, key])
    assert len(many) == 2 and np.allclose(many[1], ref, atol=atol)
