"""Noise floor of ONE full time step of BASELINE config 2 (Henon-Heiles, 64 sites, D = 64) under the reference algorithm
itself: the oracle (bit-identical to the reference on the golden runs) is re-run with its H_eff term outputs multiplied
by (1 + 2e-16 N(0,1)) -- one rounding, i.e. what any other BLAS / summation order produces.  Records how far the
autocorrelation, the energy and the Krylov iteration trace move.  Used by tests/test_gpu_bench_shapes.py to justify its
trace tolerance.      python tests/golden/make_noise_floor_c2.py   (about 2 minutes)"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle.tdvp_oracle as orc  # noqa: E402
from pytdscf_b200 import workloads  # noqa: E402


def run(eps, seed, wl):
    H = orc.MPOHamiltonian(len(wl.dims), wl.operators, wl.coupleJ)
    o = orc.TDVPOracle(H, orc.initial_mps(wl.dims, wl.bond_dim, wl.hartree, space=wl.space), integrator=wl.integrator,
                       conserve_norm=wl.conserve_norm, space=wl.space)
    rng = np.random.default_rng(seed)
    orig = orc.heff_term
    if eps:
        def noisy(L, core, R, psi):
            out = orig(L, core, R, psi)
            return out * (1 + eps * rng.standard_normal(out.shape))
        orc.heff_term = noisy
    try:
        o.propagate(wl.dt_au)
    finally:
        orc.heff_term = orig
    return o.autocorr(), o.expectation().real, list(o.trace)


def main():
    wl = workloads.by_name("c2")
    a0, e0, t0 = run(0.0, 0, wl)
    rows = []
    for seed in range(3):
        a1, e1, t1 = run(2e-16, seed, wl)
        nd = sum(1 for x, y in zip(t0, t1) if x != y)
        rows.append({"seed": seed, "autocorr_abs": abs(a1 - a0), "energy_rel": abs((e1 - e0) / e0), "trace_entries_changed": nd,
                     "max_count_change": max([abs(x[2] - y[2]) for x, y in zip(t0, t1)] + [0])})
        print(rows[-1])
    out = {"workload": wl.name, "solves_per_step": len(t0), "perturbation": 2e-16, "runs": rows,
           "autocorr_abs": max(r["autocorr_abs"] for r in rows), "energy_rel": max(r["energy_rel"] for r in rows),
           "trace_stable": all(r["trace_entries_changed"] == 0 for r in rows)}
    with open(os.path.join(HERE, "noise_floor_c2.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
