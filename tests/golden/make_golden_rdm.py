"""Reduced-density goldens: the UNMODIFIED reference's MPSCoef.get_reduced_densities (pytdscf/_mps_cls.py:1208-1678)
evaluated on the final states of existing golden runs.  Build container only.

    python tests/golden/make_golden_rdm.py        # writes tests/golden/rdm.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference  # noqa: E402

load_reference()

from pytdscf._const_cls import const  # noqa: E402
from pytdscf._mps_mpo import MPSCoefMPO  # noqa: E402
from pytdscf._site_cls import SiteCoef  # noqa: E402

from tests.golden_io import load_run  # noqa: E402

CASES = {
    # golden run -> list of remain_nleg keys (reference convention: 0 = traced, 1 = diagonal, 2 = bra and ket legs)
    "exciton_D6": [(0, 0, 0, 2), (0, 0, 0, 1), (2,), (1, 0, 2), (0, 2, 2), (2, 2, 0, 1)],
    "henon_heiles_f6": [(2,), (0, 0, 2), (0, 1, 0, 0, 0, 1), (2, 0, 0, 0, 0, 2)],
    "h2co_D16": [(0, 0, 0, 0, 0, 2), (1, 1)],
}


def main():
    const.set_runtype(jobname="golden_rdm", verbose=0)
    out = {}
    for name, keys in CASES.items():
        g = load_run(name)
        mps = MPSCoefMPO()
        mps.superblock_states = [[SiteCoef(np.array(c), "Psi" if i == 0 else "B", isite=i) for i, c in enumerate(g["final"])]]
        for k, key in enumerate(keys):
            rd = mps.get_reduced_densities(key)[0]
            out[f"{name}__{k}__key"] = np.array(key)
            out[f"{name}__{k}__rdm"] = np.array(rd)
            print(name, key, rd.shape, "trace-like", np.sum(rd) if rd.ndim == 1 else None)
        # the multi-key entry point shares environments between keys (same numbers expected)
        many = mps.get_reduced_densities(list(keys[:3]))
        for k, rd in enumerate(many):
            out[f"{name}__multi{k}__rdm"] = np.array(rd)
    # Liouville space: partial traces of the MPDO (get_partial_trace, _mps_cls.py:1438-1510)
    import importlib

    from pytdscf.model_cls import Model

    mg = importlib.import_module("tests.golden.make_golden")
    name = "liouville_spin3"
    g = load_run(name)
    basis, ops, hartree = mg.liouville_model()
    model = Model(basis, ops, bond_dim=g["bond_dim"], space="liouville")
    model.init_HartreeProduct = [hartree]
    const.set_runtype(jobname="golden_rdm_liouville", space="liouville", verbose=0)
    mps = MPSCoefMPO.alloc_random(model)
    for site, c in zip(mps.superblock_states[0], g["final"], strict=True):
        site.data = np.array(c)
    for k, key in enumerate([(2,), (0, 2), (0, 0, 2), (2, 0, 2), (1, 2), (0, 1, 2)]):
        rd = mps.get_reduced_densities(key)[0]
        out[f"{name}__{k}__key"] = np.array(key)
        out[f"{name}__{k}__rdm"] = np.array(rd)
        print(name, key, rd.shape, "trace", np.trace(rd.reshape(int(np.sqrt(rd.size)), -1)) if rd.ndim % 2 == 0 else None)
    np.savez_compressed(os.path.join(HERE, "rdm.npz"), **out)


if __name__ == "__main__":
    main()
