"""Rounding sensitivity of the reference's SITE-PARALLEL scheme (companion of make_noise_floor.py).

The product's host logic with the oracle's NumPy kernels reproduces the reference's MPSCoefParallel runs to 4e-15
(tests/test_site_parallel_cpu.py).  Here the same runs are repeated with every H_eff output multiplied by
(1 + 2e-16 N(0,1)) -- one rounding -- and the largest change of the per-step observables over 3 seeds is written to
tests/golden/noise_floor_parallel.json.  The scheme multiplies by pinv(boundary bond matrix, rcond 1e-13) and floors
singular values at 1e-4 exp(-s/1e-4), so directions at the rounding level are amplified into the observables.

    python tests/golden/make_noise_floor_parallel.py
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def worker(rank, world, port, name, tmp, eps, seed, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
    import torch

    torch.set_num_threads(1)
    import pytdscf_b200 as tb
    from oracle.oracle_engine import OracleEngine
    from pytdscf_b200 import parallel
    from tests.golden_io import load_parallel
    from tests.test_host_sweep_cpu import _build_model

    g = load_parallel(name)
    os.chdir(tmp)
    info = parallel.init_from_env("gloo")
    sim = tb.Simulator(name, _build_model(g), backend="cuda", verbose=0)
    eng = OracleEngine()
    if eps:
        rng = np.random.default_rng(1000 * seed + rank)
        orig = eng.heff_apply

        def noisy(terms, psi):
            out = orig(terms, psi)
            return out * torch.as_tensor(1 + eps * rng.standard_normal(tuple(out.shape)))

        if os.environ.get("PAR_GAUGE"):
            # exact gauge freedom of the SVD instead of rounding noise: random phases on the singular vector pairs
            eng.heff_apply = orig
            osvd, otr = eng.svd, eng.svd_truncate

            def ph(n):
                if os.environ.get("PAR_GAUGE") == "sign":
                    return rng.choice([-1.0, 1.0], n).astype(complex)
                return np.exp(2j * np.pi * rng.random(n))

            def svd(M):
                U, sv, Vh = osvd(M)
                f = ph(len(sv))
                return U * torch.as_tensor(f)[None, :], sv, Vh * torch.as_tensor(f.conj())[:, None]

            def svd_truncate(sigma, p, keepdim=False, regularize=False):
                U, S, Vh, r = otr(sigma, p, keepdim=keepdim, regularize=regularize)
                f = ph(U.shape[1])
                return U * torch.as_tensor(f)[None, :], S, Vh * torch.as_tensor(f.conj())[:, None], r

            eng.svd, eng.svd_truncate = svd, svd_truncate
        else:
            eng.heff_apply = noisy
    sim.eng = eng
    sim.rank_info = info
    sim.set_initial_mps(g["init"])
    sim.propagate(stepsize=g["dt_fs"], maxstep=g["nstep"], parallel_split_indices=g["split"], populations=False, write_files=False)
    if rank == 0:
        q.put([(complex(r["autocorr"]), float(r["energy"]), float(r["norm"])) for r in sim.history])
    parallel.finalize(info)


def run(name, eps, seed):
    import torch.multiprocessing as mp

    from tests.golden_io import load_parallel

    P = load_parallel(name)["nranks"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35000 + (os.getpid() + 17 * seed + int(eps > 0)) % 2000
    with tempfile.TemporaryDirectory() as tmp:
        procs = [ctx.Process(target=worker, args=(r, P, port, name, tmp, eps, seed, q)) for r in range(P)]
        for p in procs:
            p.start()
        hist = q.get(timeout=900)
        for p in procs:
            p.join(timeout=60)
    return np.array([h[0] for h in hist]), np.array([h[1] for h in hist]), np.array([h[2] for h in hist])


def main():
    from tests.golden_io import PAR_CASES

    path = os.path.join(HERE, f"gauge_{os.environ['PAR_GAUGE']}_floor_parallel.json" if os.environ.get("PAR_GAUGE") else "noise_floor_parallel.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    only = os.environ.get("PAR_ONLY")
    for name in PAR_CASES:
        if only and only not in name:
            continue
        a0, e0, n0 = run(name, 0.0, 0)
        fa = fe = fn = 0.0
        for seed in range(3):
            a1, e1, n1 = run(name, 2e-16, seed + 1)
            fa = max(fa, float(np.abs(a1 - a0).max()))
            fe = max(fe, float(np.abs(e1 - e0).max()))
            fn = max(fn, float(np.abs(n1 - n0).max()))
        out[name] = {"autocorr_abs": fa, "energy_abs": fe, "norm_abs": fn}
        print(name, out[name], flush=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
