"""Golden vectors of the reference's reduced densities INSIDE a site-parallel run (``MPSCoefParallel.get_reduced_densities``,
pytdscf/_mps_parallel.py:1035-1208) under the file-based mpi4py stand-in (oracle/refshim_mpi).  Build container only.

    python tests/golden/make_golden_parallel_rdm.py        # writes tests/golden/par_rdm_hh8_P2.npz, par_rdm_hh8_P4.npz

The runs are the ``par_hh8_P2`` / ``par_hh8_P4`` cases of make_golden_parallel.py (same model, split, step) with
``reduced_density=(KEYS, 1)``.  The reference writes the densities into a netCDF4 file; that package does not exist here, so
the export hook is replaced by one that records what ``get_reduced_densities`` returns on rank 0 (the propagation itself is
untouched).  Keys for which the reference raises are recorded as failures by name (see the printed summary)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
KEYS = [(0,), (5,), (2, 2), (5, 5), (3, 4), (1, 6), (3, 3, 4), (7, 7)]
NSTEP = 3


def worker(case):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim_mpi"))
    from oracle.reference_loader import load_reference

    load_reference()
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim_mpi"))
    import importlib

    mg = importlib.import_module("tests.golden.make_golden")
    mgp = importlib.import_module("tests.golden.make_golden_parallel")
    from pytdscf import properties as props
    from pytdscf._const_cls import const
    from pytdscf.model_cls import Model
    from pytdscf.simulator_cls import Simulator

    _, P, split, D, dt_fs, _ = mgp.CASES[case]
    prim, ops, vib = mg.henon_heiles_model(2000, 1.0e-3, 8, 4)
    model = Model(prim, ops, bond_dim=D)
    model.init_weight_VIBSTATE = [vib]
    rec = []

    def export(self):
        row = []
        for base_tag, rd_key in enumerate(self.rd_keys):
            try:
                row.append(self.wf.ci_coef.get_reduced_densities(base_tag, rd_key))
            except Exception as e:  # noqa: BLE001  (a failing key must not stop the other ranks' collective calls)
                row.append(f"{type(e).__name__}: {e}")
        rec.append(row)
        self.nc_row += 1

    props.Properties._create_nc_file = lambda self, reduced_density: "unused"
    props.Properties._export_reduced_density = export
    mg.RECORD["props"].clear()
    sim = Simulator(case, model, backend="numpy", verbose=0)
    sim.propagate(stepsize=dt_fs, maxstep=NSTEP, parallel_split_indices=split, populations=False, reduced_density=(KEYS, 1))
    if const.mpi_rank == 0:
        out = {"props": np.array([[t, a.real, a.imag, e.real, e.imag, n] for (t, a, e, n) in mg.RECORD["props"]])}
        for ik, key in enumerate(KEYS):
            vals = [row[ik] for row in rec]
            if all(isinstance(v, np.ndarray) for v in vals):
                out[f"rho{ik}"] = np.stack(vals)
            else:
                out[f"fail{ik}"] = np.array(str(next(v for v in vals if not isinstance(v, np.ndarray))))
        np.savez_compressed(os.path.join(os.environ["FAKE_MPI_DIR"], "rdm.npz"), **out)


def driver():
    sys.path.insert(0, ROOT)
    import importlib

    mgp = importlib.import_module("tests.golden.make_golden_parallel")
    for case in ("par_hh8_P2", "par_hh8_P4"):
        _, P, split, D, dt_fs, _ = mgp.CASES[case]
        with tempfile.TemporaryDirectory() as tmp:
            procs = []
            for r in range(P):
                env = dict(os.environ, FAKE_MPI_RANK=str(r), FAKE_MPI_SIZE=str(P), FAKE_MPI_DIR=tmp, LOGURU_LEVEL="ERROR", OPENBLAS_NUM_THREADS="1")
                procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", case], env=env, cwd=tmp,
                                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
            outs = [p.communicate(timeout=900)[0] for p in procs]
            if any(p.returncode != 0 for p in procs):
                for r, o in enumerate(outs):
                    print(f"--- rank {r} rc={procs[r].returncode}\n{o[-3000:]}")
                raise SystemExit(f"{case}: a rank failed")
            z = dict(np.load(os.path.join(tmp, "rdm.npz")))
        z["keys"] = np.array([repr(k) for k in KEYS])
        z["nstep"] = np.array(NSTEP)
        name = case.replace("par_", "par_rdm_")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **z)
        ok = [KEYS[int(k[3:])] for k in z if k.startswith("rho")]
        bad = {KEYS[int(k[4:])]: str(z[k]) for k in z if k.startswith("fail")}
        print(f"[golden-parallel-rdm] {name}: P={P} steps={NSTEP} keys ok {ok} failed in the reference {bad}")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        worker(sys.argv[2])
    else:
        driver()
