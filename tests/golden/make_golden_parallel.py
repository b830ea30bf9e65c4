"""Golden vectors of the reference's SITE-PARALLEL TDVP (MPSCoefParallel, pytdscf/_mps_parallel.py) run as P plain
processes under the file-based mpi4py stand-in (oracle/refshim_mpi).  Build container only.

    python tests/golden/make_golden_parallel.py            # driver: spawns the ranks, writes tests/golden/par_*.npz
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

CASES = {
    # name: (model, nranks, split, bond_dim, dt_fs, nstep)
    "par_exciton_P2": ("exciton", 2, [(0, 1), (2, 3)], 10, 0.05, 6),
    "par_hh8_P2": ("hh8", 2, [(0, 3), (4, 7)], 6, 0.05, 4),
    "par_hh8_P4": ("hh8", 4, [(0, 1), (2, 3), (4, 5), (6, 7)], 6, 0.05, 4),
    "par_hh8_P3": ("hh8", 3, [(0, 1, 2), (3, 4), (5, 6, 7)], 6, 0.05, 4),     # odd rank count, uneven segments
    # well-conditioned cases: an entangled random initial MPS that saturates every bond (no null space for the
    # boundary pseudo-inverse / regularised SVD to amplify), so independent implementations can agree to 1e-10
    "par_frenkel8_P2": ("frenkel8", 2, [(0, 1, 2, 3), (4, 5, 6, 7)], 4, 0.2, 10),
    "par_frenkel8_P4": ("frenkel8", 4, [(0, 1), (2, 3), (4, 5), (6, 7)], 4, 0.2, 10),
    # rank-adaptive bond dimensions INSIDE a site-parallel run (the reference's tests/test_mpi_exiciton_propagate.py[adaptive]):
    # a 7th entry = (Dmax, dD, p_proj, p_svd); the bonds grow inside the segments and across the rank boundaries
    "par_adaptive_exciton_P2": ("exciton", 2, [(0, 1), (2, 3)], 1, 0.05, 6, (8, 4, 1.0e-5, 1.0e-6)),
    "par_adaptive_hh8_P2": ("hh8", 2, [(0, 3), (4, 7)], 2, 0.1, 5, (6, 2, 1.0e-6, 1.0e-7)),
    "par_adaptive_hh8_P4": ("hh8", 4, [(0, 1), (2, 3), (4, 5), (6, 7)], 2, 0.1, 5, (6, 2, 1.0e-6, 1.0e-7)),
}


def frenkel_model(n=8, D=4, seed=77):
    """Frenkel-exciton chain  H = sum_i eps_i n_i + J sum_i (a+_i a_{i+1} + h.c.)  as a w = 4 nearest-neighbour MPO,
    started from a random MPS of full bond dimension (3-D cores through the reference's lower-level init API,
    _site_cls.py:459-472)."""
    from pytdscf.basis import Exciton
    from pytdscf.hamiltonian_cls import TensorHamiltonian
    from pytdscf.dvr_operator_cls import TensorOperator

    rng = np.random.default_rng(seed)
    prim = [Exciton(nstate=2) for _ in range(n)]
    a = np.array([[0.0, 1.0], [0.0, 0.0]], dtype=complex)
    ad, one = a.T.copy(), np.eye(2, dtype=complex)
    num = ad @ a
    eps = 0.01 + 0.004 * rng.standard_normal(n)
    J = 0.005
    cores = []
    for i in range(n):
        W = np.zeros((4, 2, 2, 4), dtype=complex)
        W[0, :, :, 0] = one
        W[1, :, :, 0] = a
        W[2, :, :, 0] = ad
        W[3, :, :, 0] = eps[i] * num
        W[3, :, :, 1] = J * ad
        W[3, :, :, 2] = J * a
        W[3, :, :, 3] = one
        if i == 0:
            W = W[3:4]
        if i == n - 1:
            W = W[:, :, :, 0:1]
        cores.append(W)
    key = tuple((i, i) for i in range(n))
    legs = tuple(j for i in range(n) for j in (i, i))
    ham = TensorHamiltonian(ndof=n, potential=[[{key: TensorOperator(mpo=cores, legs=legs)}]], backend="numpy")
    dims = [1] + [min(D, 2 ** min(i, n - i)) for i in range(1, n)] + [1]
    init = [(rng.standard_normal((dims[i], 2, dims[i + 1])) + 1j * rng.standard_normal((dims[i], 2, dims[i + 1]))) for i in range(n)]
    return prim, {"hamiltonian": ham}, init


def worker(case):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim_mpi"))
    from oracle.reference_loader import load_reference

    load_reference()
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim_mpi"))
    import importlib

    mg = importlib.import_module("tests.golden.make_golden")
    from pytdscf._const_cls import const
    from pytdscf.model_cls import Model
    from pytdscf.simulator_cls import Simulator

    model_name, P, split, D, dt_fs, nstep = CASES[case][:6]
    adaptive = CASES[case][6] if len(CASES[case]) > 6 else None
    nstep = int(os.environ.get("PAR_NSTEP", nstep))          # debugging aid: shorter runs into PAR_OUT
    if model_name == "exciton":
        prim, ops, hartree = mg.exciton_model()
        vib = None
    elif model_name == "frenkel8":
        prim, ops, hartree = frenkel_model(D=D)
        vib = None
    else:
        prim, ops, vib = mg.henon_heiles_model(2000, 1.0e-3, 8, 4)
        hartree = None
    model = Model(prim, ops, bond_dim=D)
    if hartree is not None:
        model.init_HartreeProduct = [hartree]
    if vib is not None:
        model.init_weight_VIBSTATE = [vib]
    mg.RECORD["props"].clear()
    mg.RECORD["trace"].clear()
    sim = Simulator(case, model, backend="numpy", verbose=0)
    model_dump = {}
    if int(os.environ["FAKE_MPI_RANK"]) == 0:
        # inputs, taken before the run distributes the MPO: full MPO cores and the serial initial MPS
        mpo = model.hamiltonian.mpo[0][0]
        model_dump["dims"] = np.array([len(b) for b in prim])
        model_dump["coupleJ"] = np.array(complex(model.hamiltonian.coupleJ[0][0]))
        model_dump["nkeys"] = np.array(len(mpo.operators))
        for ik, (key, cores) in enumerate(mpo.operators.items()):
            model_dump[f"key{ik}"] = np.array(repr(key))
            for ic, c in enumerate(cores):
                model_dump[f"key{ik}_core{ic}"] = np.asarray(c)
    akw = {} if adaptive is None else dict(adaptive=True, adaptive_Dmax=adaptive[0], adaptive_dD=adaptive[1],
                                           adaptive_p_proj=adaptive[2], adaptive_p_svd=adaptive[3])
    # PAR_PERTURB (adaptive cases, second run): the time step scaled by 1 + 1e-14 -- a rounding-level perturbation that shows how
    # far the reference reproduces ITSELF (the regularised QR / SVD of the boundary update floor singular values of 1e-8 ... 1e-19
    # to 1e-4 along singular vectors that are rounding noise, so the result is defined only up to that noise)
    dt_fs = dt_fs * (1.0 + float(os.environ.get("PAR_PERTURB", "0")))
    ener, wf = sim.propagate(stepsize=dt_fs, maxstep=nstep, parallel_split_indices=split, populations=False, **akw)
    rank = const.mpi_rank
    if rank == 0:
        from pytdscf._mps_mpo import MPSCoefMPO   # the serial initial MPS that rank 0 canonicalises and distributes

        for i, s in enumerate(MPSCoefMPO.alloc_random(model).superblock_states[0]):
            model_dump[f"init{i}"] = np.array(s.data)
    out = {"props": np.array([[t, a.real, a.imag, e.real, e.imag, n] for (t, a, e, n) in mg.RECORD["props"]]) if rank == 0 else np.zeros(0),
           "final_energy": np.array(complex(ener) if ener is not None else np.nan)}
    out.update(model_dump)
    out["trace"] = np.array(mg.RECORD["trace"], dtype=np.int64).reshape(-1, 3)    # this rank's Krylov solves: kind, local site, niter
    mps = wf.ci_coef
    for i, s in enumerate(mps.superblock_states[0]):
        out[f"site{i}"] = np.array(s.data)
        out[f"gauge{i}"] = np.array(s.gauge)
    if hasattr(mps, "joint_sigvec") and rank != const.mpi_size - 1:
        out["joint_sigvec"] = np.array(mps.joint_sigvec)
        out["joint_sigvec_not_pinv"] = np.array(mps.joint_sigvec_not_pinv)
    np.savez_compressed(os.path.join(os.environ["FAKE_MPI_DIR"], f"rank{rank}.npz"), **out)


def driver():
    sys.path.insert(0, ROOT)
    only = os.environ.get("PAR_ONLY")
    for case, spec in CASES.items():
        model_name, P, split, D, dt_fs, nstep = spec[:6]
        if only and only not in case:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            procs = []
            for r in range(P):
                env = dict(os.environ, FAKE_MPI_RANK=str(r), FAKE_MPI_SIZE=str(P), FAKE_MPI_DIR=tmp, LOGURU_LEVEL="ERROR",
                           OPENBLAS_NUM_THREADS="1")
                procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", case], env=env, cwd=tmp,
                                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
            outs = [p.communicate(timeout=900)[0] for p in procs]
            if any(p.returncode != 0 for p in procs):
                for r, o in enumerate(outs):
                    print(f"--- rank {r} rc={procs[r].returncode}\n{o[-3000:]}")
                raise SystemExit(f"{case}: a rank failed")
            noise = None
            if len(spec) > 6:      # the same run again under a rounding-level perturbation: the reference's own reproducibility
                with tempfile.TemporaryDirectory() as tmp2:
                    procs2 = []
                    for r in range(P):
                        env = dict(os.environ, FAKE_MPI_RANK=str(r), FAKE_MPI_SIZE=str(P), FAKE_MPI_DIR=tmp2, LOGURU_LEVEL="ERROR",
                                   OPENBLAS_NUM_THREADS="1", PAR_PERTURB="1e-14")
                        procs2.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", case], env=env, cwd=tmp2,
                                                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
                    for p2 in procs2:
                        p2.communicate(timeout=900)
                    if any(p2.returncode != 0 for p2 in procs2):
                        raise SystemExit(f"{case}: a rank of the perturbed run failed")
                    noise = np.load(os.path.join(tmp2, "rank0.npz"))["props"]
            merged = {"nranks": np.array(P), "split": np.array([(seg[0], seg[-1]) for seg in split]), "bond_dim": np.array(D), "dt_fs": np.array(dt_fs),
                      "nstep": np.array(nstep), "model": np.array(model_name)}
            if len(spec) > 6:
                merged["adaptive"] = np.array(spec[6], dtype=float)
                merged["props_perturbed"] = noise
            for r in range(P):
                z = np.load(os.path.join(tmp, f"rank{r}.npz"))
                for k in z.files:
                    if k == "dims" or k == "coupleJ" or k == "nkeys" or k.startswith("key") or k.startswith("init"):
                        merged[k] = z[k]
                    else:
                        merged[f"r{r}_{k}"] = z[k]
            np.savez_compressed(os.path.join(os.environ.get("PAR_OUT", HERE), case + ".npz"), **merged)
            pr = merged["r0_props"]
            print(f"[golden-parallel] {case}: P={P} steps={nstep} E0={pr[0, 3]:.12f} E_last={pr[-1, 3]:.12f} "
                  f"norm_last={pr[-1, 5]:.10f} autocorr_last={pr[-1, 1]:+.8f}{pr[-1, 2]:+.8f}j")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        worker(sys.argv[2])
    else:
        driver()
