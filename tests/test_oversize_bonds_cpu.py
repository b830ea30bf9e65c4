"""CPU host-logic test: explicit site tensors whose bonds exceed the reference's bond-dimension rule
(``LatticeInfo.get_bond_dim``, pytdscf/_mps_cls.py:2616-2631) are compressed exactly at upload, which is what the
reference's economic QR does implicitly at its first sweep (``gauge_trf``, pytdscf/_site_cls.py:138-292); NumPy kernels of
the oracle injected in place of the CUDA engine."""
import numpy as np
import pytest

from oracle.oracle_engine import OracleEngine
from pytdscf_b200._mps_cuda import MPSCoefCuda, oversize_bonds
from tests.device_numerics_engine import DeviceNumericsEngine

ENGINES = [OracleEngine, DeviceNumericsEngine]     # LAPACK's SVD conventions / the device's (tests/device_numerics_engine.py)


def dense(cores):
    t = np.asarray(cores[0])
    for c in cores[1:]:
        t = np.tensordot(t, np.asarray(c), axes=(-1, 0))
    return t.reshape(-1)


def rand_chain(rng, dims, bonds):
    b = [1, *bonds, 1]
    return [rng.standard_normal((b[i], d, b[i + 1])) + 1j * rng.standard_normal((b[i], d, b[i + 1])) for i, d in enumerate(dims)]


@pytest.mark.parametrize("dims,bonds", [([2, 4, 3, 2], [4, 9, 5]),       # wide first site, oversize middle bond, tall last site
                                         ([2, 2, 2, 2, 2], [7, 7, 7, 7]),  # every bond far beyond 2^k
                                         ([3, 3, 3], [3, 3])])             # already within the rule: untouched
@pytest.mark.parametrize("engine", ENGINES)
def test_oversize_bonds_are_compressed_exactly(dims, bonds, engine):
    rng = np.random.default_rng(sum(bonds))
    cores = rand_chain(rng, dims, bonds)
    ref = dense(cores)
    eng = engine()
    before = oversize_bonds(MPSCoefCuda(eng, [eng.to_device(c) for c in cores]).sites)
    mps = MPSCoefCuda.from_user_cores(eng, cores)
    got = [np.asarray(s.data) for s in mps.sites]
    if not before:
        assert all(np.array_equal(g, c) for g, c in zip(got, cores, strict=True))
        return
    assert not oversize_bonds(mps.sites)
    assert np.abs(dense(got) - ref).max() < 1e-12 * np.abs(ref).max()
    assert [s.gauge for s in mps.sites] == ["Psi"] + ["B"] * (len(dims) - 1)
    for g in got[1:]:
        m = g.reshape(g.shape[0], -1)
        assert np.abs(m @ m.conj().T - np.eye(m.shape[0])).max() < 1e-12
    # every bond now obeys D_r <= D_l d and D_l <= d D_r on both sides
    for g in got:
        assert g.shape[2] <= g.shape[0] * g.shape[1] and g.shape[0] <= g.shape[1] * g.shape[2]


def test_propagation_from_oversize_initial_mps_runs(tmp_path):
    """Simulator.set_initial_mps with a bond the QR gauge shift could not factor: the run starts from the compressed chain and
    gives the energies of the same state uploaded within the rule."""
    import os

    import pytdscf_b200 as tb
    from tests.golden_io import load_run
    from tests.test_gpu_propagation import build_model

    g = load_run("exciton_D6")
    rng = np.random.default_rng(11)
    dims = list(g["dims"])
    bonds = [min(6, int(np.prod(dims[:i + 1])), int(np.prod(dims[i + 1:]))) for i in range(len(dims) - 1)]
    cores = rand_chain(rng, dims, bonds)
    for i in range(len(cores) - 1, 0, -1):       # right-canonical form with the centre on site 0 (the default gauge labels)
        a, d, b = cores[i].shape
        q, r = np.linalg.qr(cores[i].reshape(a, d * b).conj().T)
        cores[i] = q.conj().T.reshape(-1, d, b)
        cores[i - 1] = np.tensordot(cores[i - 1], r.conj().T, axes=(2, 0))
    cores[0] = cores[0] / np.linalg.norm(cores[0])
    # blow the first bond up beyond d_0 with a random isometric insertion: same state, oversize bond
    D = cores[0].shape[2]
    q, _ = np.linalg.qr(rng.standard_normal((D + 3, D)) + 1j * rng.standard_normal((D + 3, D)))
    fat = [np.tensordot(cores[0], q.conj().T, axes=(2, 0)), np.tensordot(q, cores[1], axes=(1, 0)), *cores[2:]]
    assert np.abs(dense(fat) - dense(cores)).max() < 1e-13
    os.chdir(tmp_path)
    out = []
    for tag, cs in (("fat", fat), ("slim", cores)):
        sim = tb.Simulator("ov_" + tag, build_model(g), backend="cuda", verbose=0)
        sim.eng = OracleEngine()
        sim.set_initial_mps(cs)
        sim.propagate(stepsize=0.02 * g["dt_au"] * tb.units.au_in_fs, maxstep=2, thresh_sil=g["thresh_sil"], autocorr=False, populations=False)
        out.append([rec["energy"] for rec in sim.history])
    assert np.allclose(out[0], out[1], rtol=1e-9, atol=0)
