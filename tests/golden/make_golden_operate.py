"""Golden vectors of the reference's ``Simulator.operate`` (|Psi'> = O|Psi> / |O|Psi>| by fitting sweeps, pytdscf/simulator_cls.py:286-330,
wavefunction.py:303-351, _mps_cls.py:421-450, 718-796, 2733-2778) on the MPO path.  Build container only.

    python tests/golden/make_golden_operate.py            # writes tests/golden/operate.npz

Two cases on the exciton model of tests/test_exiciton_propagate.py (3 HO-DVR modes + 2-level site):
  prod   O applied to the initial Hartree product (bond dimension 4, zero-padded),
  prop   O applied to the state after three propagation steps (entangled; the reference's relax -> operate -> propagate workflow
         loads the previous run's wf_<job><ext>.pkl the same way).
O = c + sum-of-products MPO of bond dimension 2 with diagonal cores on the modes and a full core on the exciton site (seeded random
numbers: the operator is neither Hermitian nor normalised), c = 0.3 - 0.1i (``coupleJ``)."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def operator_cores(rng, nprim=8):
    def c(*shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    return [c(1, nprim, 2), c(2, nprim, 2), c(2, nprim, 2), c(2, 2, 2, 1)]


def main():
    from oracle.reference_loader import load_reference

    load_reference()
    import importlib

    mg = importlib.import_module("tests.golden.make_golden")
    from pytdscf.dvr_operator_cls import TensorOperator
    from pytdscf.hamiltonian_cls import TensorHamiltonian
    from pytdscf.model_cls import Model
    from pytdscf.simulator_cls import Simulator

    rng = np.random.default_rng(2024)
    prim, ops, hartree = mg.exciton_model()
    cores = operator_cores(rng)
    coupleJ = 0.3 - 0.1j
    out = {"coupleJ": np.array(coupleJ), "dims": np.array([len(b) for b in prim])}
    for i, c in enumerate(cores):
        out[f"O{i}"] = c
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for tag, nprop in (("prod", 0), ("prop", 3)):
                job = "operate_" + tag
                model = Model(prim, ops, bond_dim=4)
                model.init_HartreeProduct = [hartree]
                sim = Simulator(job, model, backend="numpy", verbose=0)
                ener, wf = sim.propagate(stepsize=0.1, maxstep=nprop, savefile_ext="_p", autocorr=False, populations=False)
                for i, s in enumerate(wf.ci_coef.superblock_states[0]):
                    out[f"{tag}_init{i}"] = np.array(s.data)
                out[f"{tag}_init_gauges"] = np.array([s.gauge for s in wf.ci_coef.superblock_states[0]])
                pot = {(0, 1, 2, (3, 3)): TensorOperator(mpo=[c.copy() for c in cores], legs=(0, 1, 2, 3, 3)), (): coupleJ}
                op = TensorHamiltonian(ndof=4, potential=[[pot]], kinetic=None, backend="numpy")
                model_o = Model(prim, {"hamiltonian": op}, bond_dim=4)
                model_o.init_HartreeProduct = [hartree]
                sim_o = Simulator(job, model_o, backend="numpy", verbose=0)
                norm, wf_o = sim_o.operate(restart=True, loadfile_ext="_p", maxstep=10)
                out[f"{tag}_norm"] = np.array(float(norm))
                for i, s in enumerate(wf_o.ci_coef.superblock_states[0]):
                    out[f"{tag}_final{i}"] = np.array(s.data)
                out[f"{tag}_final_gauges"] = np.array([s.gauge for s in wf_o.ci_coef.superblock_states[0]])
                print(f"[golden-operate] {tag}: norm = {norm!r}, gauges {[s.gauge for s in wf_o.ci_coef.superblock_states[0]]}, "
                      f"shapes {[s.data.shape for s in wf_o.ci_coef.superblock_states[0]]}")
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "operate.npz"), **out)


if __name__ == "__main__":
    main()
