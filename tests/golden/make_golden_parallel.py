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
}


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

    model_name, P, split, D, dt_fs, nstep = CASES[case]
    if model_name == "exciton":
        prim, ops, hartree = mg.exciton_model()
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
    ener, wf = sim.propagate(stepsize=dt_fs, maxstep=nstep, parallel_split_indices=split, populations=False)
    rank = const.mpi_rank
    if rank == 0:
        from pytdscf._mps_mpo import MPSCoefMPO   # the serial initial MPS that rank 0 canonicalises and distributes

        for i, s in enumerate(MPSCoefMPO.alloc_random(model).superblock_states[0]):
            model_dump[f"init{i}"] = np.array(s.data)
    out = {"props": np.array([[t, a.real, a.imag, e.real, e.imag, n] for (t, a, e, n) in mg.RECORD["props"]]) if rank == 0 else np.zeros(0),
           "final_energy": np.array(complex(ener) if ener is not None else np.nan)}
    out.update(model_dump)
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
    for case, (model_name, P, split, D, dt_fs, nstep) in CASES.items():
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
            merged = {"nranks": np.array(P), "split": np.array(split), "bond_dim": np.array(D), "dt_fs": np.array(dt_fs),
                      "nstep": np.array(nstep), "model": np.array(model_name)}
            for r in range(P):
                z = np.load(os.path.join(tmp, f"rank{r}.npz"))
                for k in z.files:
                    if k == "dims" or k == "coupleJ" or k == "nkeys" or k.startswith("key") or k.startswith("init"):
                        merged[k] = z[k]
                    else:
                        merged[f"r{r}_{k}"] = z[k]
            np.savez_compressed(os.path.join(HERE, case + ".npz"), **merged)
            pr = merged["r0_props"]
            print(f"[golden-parallel] {case}: P={P} steps={nstep} E0={pr[0, 3]:.12f} E_last={pr[-1, 3]:.12f} "
                  f"norm_last={pr[-1, 5]:.10f} autocorr_last={pr[-1, 1]:+.8f}{pr[-1, 2]:+.8f}j")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        worker(sys.argv[2])
    else:
        driver()
