"""``Simulator.operate`` (|Psi'> = O|Psi> / |O|Psi>|, reference simulator_cls.py:286-330, _mps_cls.py:421-450, 718-796) -- CPU host
logic with the oracle's kernels against the unmodified reference's output (tests/golden/operate.npz, make_golden_operate.py) and
against the literal its own test pins."""
import os

import numpy as np
import pytest

from oracle.oracle_engine import OracleEngine
from tests.device_numerics_engine import DeviceNumericsEngine
from tests.golden_io import GOLDEN_DIR


def dense(cores):
    t = np.asarray(cores[0])
    for c in cores[1:]:
        t = np.tensordot(t, np.asarray(c), axes=(-1, 0))
    return t.reshape(-1)


def operator_model(z):
    import pytdscf_b200 as tb

    basis = [tb.Exciton(nstate=int(d)) for d in z["dims"]]
    cores = [z[f"O{i}"] for i in range(4)]
    pot = {(0, 1, 2, (3, 3)): tb.TensorOperator(mpo=cores, legs=(0, 1, 2, 3, 3)), (): complex(z["coupleJ"])}
    op = tb.TensorHamiltonian(ndof=4, potential=[[pot]], backend="cuda")
    return tb.Model(basis, {"hamiltonian": op}, bond_dim=4), cores


@pytest.mark.parametrize("engine", [OracleEngine, DeviceNumericsEngine])
@pytest.mark.parametrize("tag", ["prod", "prop"])
def test_operate_matches_reference(tag, engine, tmp_path):
    import pytdscf_b200 as tb
    from pytdscf_b200.mpo_tools import mpo_to_dense

    z = dict(np.load(os.path.join(GOLDEN_DIR, "operate.npz")))
    model, cores = operator_model(z)
    os.chdir(tmp_path)
    sim = tb.Simulator("operate_" + tag, model, backend="cuda", verbose=0)
    sim.eng = engine()
    init = [z[f"{tag}_init{i}"] for i in range(4)]
    sim.set_initial_mps(init, [str(g) for g in z[f"{tag}_init_gauges"]])
    norm, wf = sim.operate(maxstep=10)
    assert abs(norm - float(z[f"{tag}_norm"])) < 1e-10 * float(z[f"{tag}_norm"])
    got = wf.ci_coef.to_numpy()
    final = [z[f"{tag}_final{i}"] for i in range(4)]
    assert [c.shape for c in got] == [c.shape for c in final]
    assert [s.gauge for s in wf.ci_coef.sites] == [str(g) for g in z[f"{tag}_final_gauges"]]
    np.testing.assert_allclose(dense(got), dense(final), rtol=0, atol=1e-10)
    # and against the definition: O|Psi> with the dense operator, for the state the fit can represent exactly or nearly so
    O = mpo_to_dense(cores) + complex(z["coupleJ"]) * np.eye(dense(init).size)
    exact = O @ dense(init)
    fit = norm * dense(got)
    assert abs(np.linalg.norm(exact) - norm) < (1e-9 if tag == "prod" else 5e-2) * np.linalg.norm(exact)
    assert np.linalg.norm(fit - exact) < (1e-8 if tag == "prod" else 0.35) * np.linalg.norm(exact)
    assert os.path.exists(f"wf_operate_{tag}_operate.pkl") and os.path.exists(f"operate_{tag}_operate/main.log")


def test_relax_operate_workflow_reproduces_the_reference_literal(tmp_path):
    """The reference's tests/test_harmonic_dvr_func_full_mpssm_jax.py:59-100: the relaxed ground state of three harmonic modes,
    then ``operate(restart=True, maxstep=5)`` with O = 1 + 0.1 (q1 + q2 + q3) gives the pinned norm 1.6490051381599562.  The
    separable dipole surface needs no MPO builder (three one-site diagonal keys + the scalar term)."""
    import pytdscf_b200 as tb

    os.chdir(tmp_path)
    freqs = [1500, 2000, 2500]
    prim = [tb.HarmonicOscillator(5, w, 0.0) for w in freqs]
    pot, dms = {}, {(): 1.0}
    for i, (p, w) in enumerate(zip(prim, freqs)):
        q = np.array(p.get_grids())
        pot[(i,)] = tb.TensorOperator(mpo=[((w / tb.units.au_in_cm1) ** 2 / 2 * q**2).reshape(1, 5, 1)], legs=(i,))
        dms[(i,)] = tb.TensorOperator(mpo=[(0.1 * q).reshape(1, 5, 1)], legs=(i,))
    ham = tb.TensorHamiltonian(ndof=3, potential=[[pot]], kinetic=[[tb.construct_kinetic_operator(dvr_prims=prim)]], backend="cuda")
    model = tb.Model(tb.BasInfo([prim]), {"hamiltonian": ham})
    model.m_aux_max = 4
    sim = tb.Simulator("harmonic_dvr", model, backend="cuda", verbose=0)
    sim.eng = OracleEngine()
    sim.relax(maxstep=3, stepsize=0.1)                                    # writes wf_harmonic_dvr_gs.pkl
    model_o = tb.Model(tb.BasInfo([prim]), {"hamiltonian": tb.TensorHamiltonian(ndof=3, potential=[[dms]], kinetic=None, backend="cuda")})
    model_o.m_aux_max = 4
    sim_o = tb.Simulator("harmonic_dvr", model_o, backend="cuda", verbose=0)
    sim_o.eng = OracleEngine()
    norm, wf = sim_o.operate(restart=True, maxstep=5)
    assert norm == pytest.approx(1.6490051381599562)
    # ... and the spectra workflow continues from the operated state
    sim_p = tb.Simulator("harmonic_dvr", model, backend="cuda", verbose=0)
    sim_p.eng = OracleEngine()
    ener, wf = sim_p.propagate(maxstep=3, stepsize=0.1, restart=True)
    assert ener == pytest.approx(0.019185297685193108)                    # tests/test_harmonic_dvr_func_full_mpssm_jax.py:147
