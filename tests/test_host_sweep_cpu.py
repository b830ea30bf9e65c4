"""CPU tests of the PRODUCT's host logic (pytdscf_b200/_mps_cuda.py, Simulator loop, multi-process plumbing) with the
oracle's NumPy kernels injected in place of the CUDA engine (oracle/oracle_engine.py, test infrastructure only)."""
import os

import numpy as np
import pytest
import torch

from oracle.oracle_engine import OracleEngine
from tests.golden_io import GATE_CASES, RUN_CASES, load_run


def _build_model(g):
    import pytdscf_b200 as tb

    basis = [tb.Exciton(nstate=d) for d in g["dims"]]
    pot = {key: tb.TensorOperator(mpo=[np.asarray(c) for c in cores]) for key, cores in g["operators"].items()}
    ham = tb.TensorHamiltonian(ndof=len(basis), potential=[[pot]], backend="cuda")
    gate = None
    if g.get("gates"):
        gp = {}
        for site, U in g["gates"].items():
            if U.ndim == 1:
                gp[(site,)] = tb.TensorOperator(mpo=[U.reshape(1, -1, 1)], legs=(site,))
            else:
                gp[((site, site),)] = tb.TensorOperator(mpo=[U.reshape(1, U.shape[0], U.shape[1], 1)], legs=(site, site))
        gate = tb.TensorHamiltonian(ndof=len(basis), potential=[[gp]], backend="cuda")
    return tb.Model(basis, {"hamiltonian": ham}, bond_dim=g["bond_dim"], space=g["space"], one_gate_to_apply=gate)


def _run(g, tmp_path, relax=None):
    import pytdscf_b200 as tb

    model = _build_model(g)
    os.chdir(tmp_path)
    sim = tb.Simulator(g["name"] + "_cpu", model, backend="cuda")
    sim.eng = OracleEngine()  # inject: Simulator would otherwise insist on a CUDA device
    sim.set_initial_mps(g["init"])
    hil = g["space"] == "hilbert"
    if relax is None:
        ener, wf = sim.propagate(stepsize=g["dt_au"] * tb.units.au_in_fs, maxstep=g["nstep"], thresh_sil=g["thresh_sil"],
                                 integrator=g["integrator"], conserve_norm=g["conserve_norm"], energy=hil, autocorr=hil,
                                 norm=hil, populations=hil, record_trace=True)
    else:
        ener, wf = sim.relax(stepsize=g["dt_au"] * tb.units.au_in_fs, maxstep=g["nstep"], improved=(relax == "improved"),
                             record_trace=True)
    return sim, ener, wf


@pytest.mark.parametrize("name", RUN_CASES + GATE_CASES)
def test_host_sweep_logic_reproduces_reference(name, tmp_path):
    """Same kernels as the oracle => the product's bookkeeping must reproduce the reference run exactly."""
    g = load_run(name)
    sim, ener, wf = _run(g, tmp_path)
    assert (np.array(wf.ci_coef.trace) == g["trace"]).all()
    if g["space"] == "hilbert":
        for rec, row in zip(sim.history, g["props"], strict=True):
            assert abs(rec["autocorr"] - complex(row[1], row[2])) < 1e-13
            assert abs(rec["energy"] - row[3]) < 1e-13 * max(1.0, abs(row[3]))
            assert abs(rec["norm"] - row[5]) < 1e-13
    for c, r in zip(wf.ci_coef.to_numpy(), g["final"], strict=True):
        np.testing.assert_allclose(c, r, rtol=0, atol=1e-12)


@pytest.mark.parametrize("name,mode", [("relax_improved_hh4", "improved"), ("relax_imag_hh4", "imag")])
def test_host_relaxation_logic(name, mode, tmp_path):
    g = load_run(name)
    sim, ener, wf = _run(g, tmp_path, relax=mode)
    assert (np.array(wf.ci_coef.trace) == g["trace"]).all()
    for rec, row in zip(sim.history, g["props"], strict=True):
        assert abs(rec["energy"] - row[3]) < 1e-12 * abs(row[3])
    assert abs(ener - g["final_energy"].real) < 1e-12 * abs(g["final_energy"].real)


def test_device_alloc_random_logic_on_cpu():
    from pytdscf_b200._mps_cuda import MPSCoefCuda

    g = load_run("exciton_D6")
    model = _build_model(g)
    model.init_HartreeProduct = [[h for h in g["hartree"]]]
    mps = MPSCoefCuda.alloc_random(OracleEngine(), model)
    for c, r in zip(mps.to_numpy(), g["init"], strict=True):
        np.testing.assert_allclose(c, r, rtol=0, atol=1e-15)
    assert mps.bonddim() == [s.shape[2] for s in g["init"][:-1]]


def test_checkpoint_roundtrip(tmp_path):
    g = load_run("henon_heiles_f2")
    sim, ener, wf = _run(g, tmp_path)
    path = sim.save_wavefunction(wf, "_ck")
    assert os.path.exists(path)
    wf2 = sim.load_wavefunction("_ck")
    for a, b in zip(wf.ci_coef.to_numpy(), wf2.ci_coef.to_numpy(), strict=True):
        assert (a == b).all()
    assert [s.gauge for s in wf2.ci_coef.sites] == [s.gauge for s in wf.ci_coef.sites]


def _replica_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import time

    from pytdscf_b200 import parallel
    from pytdscf_b200._const_cls import RunConfig
    from pytdscf_b200._mps_cuda import DeviceMPO, MPSCoefCuda

    info = parallel.init_from_env("gloo")
    g = load_run("henon_heiles_f2")
    eng = OracleEngine()
    model = _build_model(g)
    H = DeviceMPO(eng, model.hamiltonian)
    mps = MPSCoefCuda(eng, [eng.to_device(c) for c in g["init"]])
    cfg = RunConfig()
    parallel.barrier(info)
    t0 = time.perf_counter()
    for _ in range(g["nstep"]):
        mps.propagate(g["dt_au"], H, cfg)
    sec = time.perf_counter() - t0 + 0.01 * rank  # make the ranks' clocks differ
    slowest = parallel.max_over_ranks(info, sec)
    total = parallel.aggregate_throughput(info, 2 * g["nstep"], sec)
    e = mps.expectation(H).real
    q.put((rank, sec, slowest, total, e))
    parallel.finalize(info)


def test_replicas_world_size_2_gloo():
    """N > 1 path of round 1: independent replicas, control-plane reductions only (MAX of times, SUM of work)."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_replica_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = load_run("henon_heiles_f2")
    secs = [r[1] for r in res]
    for rank, sec, slowest, total, e in res:
        assert slowest == pytest.approx(max(secs))
        assert total == pytest.approx(2 * (2 * g["nstep"]) / max(secs))
        assert abs(e - g["final_energy"].real) < 1e-9  # energy is conserved; every replica ran the same physics


def test_reduced_densities_host_logic():
    """Index bookkeeping of the device reduced-density contraction against the reference's get_reduced_densities."""
    from pytdscf_b200._mps_cuda import MPSCoefCuda
    from tests.rdm_cases import check_rdms

    check_rdms(OracleEngine(), load_run, MPSCoefCuda, atol=1e-14)


def test_simulator_reduced_density_argument(tmp_path):
    import pytdscf_b200 as tb

    g = load_run("exciton_D6")
    model = _build_model(g)
    os.chdir(tmp_path)
    sim = tb.Simulator("rd_cpu", model, backend="cuda")
    sim.eng = OracleEngine()
    sim.set_initial_mps(g["init"])
    sim.propagate(stepsize=g["dt_au"] * tb.units.au_in_fs, maxstep=3, reduced_density=([(3, 3), (0,), (1, 2)], 2))
    recs = [r for r in sim.history if "reduced_densities" in r]
    assert len(recs) == 2
    rho = recs[0]["reduced_densities"]
    assert rho[(3, 3)].shape == (2, 2) and rho[(0,)].shape == (8,) and rho[(1, 2)].shape == (8, 8)
    assert abs(np.trace(rho[(3, 3)]) - 1.0) < 1e-12 and abs(rho[(0,)].sum() - 1.0) < 1e-12
    z = np.load(os.path.join("rd_cpu_prop", "reduced_density.npz"))
    assert z["rho_3_3"].shape == (2, 2, 2) and len(z["time_au"]) == 2
    # <job>/reduced_density.nc in the reference's layout (pytdscf/properties.py:160-213): step / Q<site> dimensions, `time`,
    # rho_<key>_<state>; NetCDF-3 with a trailing (real, imag) dimension instead of NETCDF4's compound type
    from scipy.io import netcdf_file

    with netcdf_file(os.path.join("rd_cpu_prop", "reduced_density.nc"), "r", mmap=False) as f:
        assert f.dimensions["step"] is None and f.dimensions["Q3"] == 2 and f.dimensions["Q1"] == 8 and f.dimensions["state"] == 1
        assert set(f.variables) == {"time", "rho_(3, 3)_0", "rho_(0,)_0", "rho_(1, 2)_0"}
        assert f.variables["rho_(1, 2)_0"].dimensions == ("step", "Q1", "Q2", "complex")
        assert np.allclose(f.variables["time"][:], [r["time_au"] * tb.units.au_in_fs for r in recs])
        for key, name in (((3, 3), "rho_(3, 3)_0"), ((0,), "rho_(0,)_0"), ((1, 2), "rho_(1, 2)_0")):
            v = f.variables[name][:]
            assert np.array_equal(v[..., 0] + 1j * v[..., 1], np.stack([r["reduced_densities"][key] for r in recs]))


@pytest.mark.parametrize("reorder", [False, True])
@pytest.mark.parametrize("name", ["exciton_D6", "henon_heiles_f6", "h2co_D16", "liouville_spin3"])
def test_identity_channel_analysis_matches_the_environments(name, reorder):
    """``identity_channels`` (MPO structure only) must name exactly channels whose environment block, contracted the
    reference's way from canonical tensors, IS the unit matrix -- that is what lets H_eff / K_eff copy instead of multiply."""
    from pytdscf_b200._mps_cuda import DeviceMPO, MPSCoefCuda

    g = load_run(name)
    eng = OracleEngine()
    eng.reorder_mpo_channels = reorder
    model = _build_model(g)
    H = DeviceMPO(eng, model.hamiltonian)
    if reorder:   # flagged channels sit at the ends of the bond, and the operator is unchanged
        for terms in H.calc_point:
            for term in terms:
                assert term.core.l_id in (-1, 0) and term.core.r_id in (-1, term.core.wr - 1)
        H0 = DeviceMPO(OracleEngine(), model.hamiltonian)
        m0 = MPSCoefCuda(eng, [eng.to_device(c) for c in g["final"]])
        assert abs(m0.expectation(H) - m0.expectation(H0)) < 1e-13 * max(1.0, abs(m0.expectation(H0)))
    mps = MPSCoefCuda(eng, [eng.to_device(c) for c in g["final"]])      # Psi B B ... B
    n = mps.nsite
    right = mps.construct_op_sites(n - 1, 0, H)                             # right[k]: block right of site n-1-k ... built from B
    found = 0
    for p in range(n - 1):
        blocks = right[n - 1 - p]                                           # environment of site p (sites p+1 .. n-1)
        for term in H.calc_point[p]:
            if term.core.r_id >= 0 and term.key in blocks:
                E = blocks[term.key].numpy()
                np.testing.assert_allclose(E[:, term.core.r_id, :], np.eye(E.shape[0]), atol=1e-12)
                found += 1
    # left blocks from A tensors: shift the centre to the right end first
    cfg_sites = [s.data for s in mps.sites]
    for i in range(n - 1):
        A, sig = eng.qr_shift("A", cfg_sites[i])
        cfg_sites[i] = A
        cfg_sites[i + 1] = eng.absorb("A", sig, cfg_sites[i + 1])
    mpsA = MPSCoefCuda(eng, cfg_sites, ["A"] * (n - 1) + ["Psi"])
    left = mpsA.construct_op_sites(0, n - 1, H)                             # left[p]: block of sites 0 .. p-1
    for p in range(1, n):
        for term in H.calc_point[p]:
            if term.core.l_id >= 0 and term.key in left[p]:
                E = left[p][term.key].numpy()
                np.testing.assert_allclose(E[:, term.core.l_id, :], np.eye(E.shape[0]), atol=1e-12)
                found += 1
        for (key, b), (lid, rid) in H.bond_ids.items():
            if b == p and key in left[p] and lid >= 0:
                np.testing.assert_allclose(left[p][key].numpy()[:, lid, :], np.eye(left[p][key].shape[0]), atol=1e-12)
    assert found > 0, "no identity channel detected in an MPO that has prefix/suffix channels"


def _kraus_model(g):
    import pytdscf_b200 as tb

    basis = [tb.Exciton(nstate=d) for d in g["dims"]]
    pot = {key: tb.TensorOperator(mpo=[np.asarray(c) for c in cores]) for key, cores in g["operators"].items()}
    ham = tb.TensorHamiltonian(ndof=len(basis), potential=[[pot]], backend="cuda")
    return tb.Model(basis, {"hamiltonian": ham}, bond_dim=g["bond_dim"], kraus_op=g["kraus"])


def kraus_observables(sim, wf, g):
    """Quantities that do not depend on the phases of the ancilla basis the SVD picks: per-step energy and norm, and the
    system part of the Kraus site's reduced density (ancilla traced out)."""
    K = g["kraus_K"]
    rd = wf.get_reduced_densities((0, 2))[0]
    if len(next(iter(g["kraus"]))) == 2:      # two-site form: site 1 is the bare system site
        rho, ref = rd, g["rdm_site1"]
    else:
        d = rd.shape[0] // K
        rho = np.einsum("akbk->ab", rd.reshape(d, K, d, K))
        ref = np.einsum("akbk->ab", g["rdm_site1"].reshape(d, K, d, K))
    return [(r["energy"], r["norm"]) for r in sim.history], rho, ref


@pytest.mark.parametrize("engine", ["lapack", "device"])
@pytest.mark.parametrize("case", ["kraus_spin4", "kraus2_spin4"])
def test_host_kraus_logic(case, engine, tmp_path):
    """Kraus maps between the half sweeps: with the oracle's kernels (same LAPACK SVD) the product reproduces the
    reference run, Krylov trace included -- and also with the SVD following the device's conventions
    (tests/device_numerics_engine.py): the observables are invariant under the ancilla gauge the SVD leaves open."""
    import pytdscf_b200 as tb
    from tests.device_numerics_engine import DeviceNumericsEngine

    g = load_run(case)
    os.chdir(tmp_path)
    sim = tb.Simulator("kraus_cpu", _kraus_model(g), backend="cuda")
    sim.eng = OracleEngine() if engine == "lapack" else DeviceNumericsEngine()
    sim.set_initial_mps(g["init"])
    ener, wf = sim.propagate(stepsize=g["dt_au"] * tb.units.au_in_fs, maxstep=g["nstep"], autocorr=False, populations=False,
                             conserve_norm=False, record_trace=True)
    assert (np.array(wf.ci_coef.trace) == g["trace"]).all()
    obs, rho, ref = kraus_observables(sim, wf, g)
    for (e, n), row in zip(obs, g["props"], strict=True):
        assert abs(e - row[3]) < 1e-12 and abs(n - row[5]) < 1e-12
    np.testing.assert_allclose(rho, ref, atol=1e-12)
    assert [s.shape for s in wf.ci_coef.sites] == [c.shape for c in g["final"]]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_identity_channels_on_random_sum_of_products(seed):
    """Random sum-of-products operators -> sop_to_mpo -> DeviceMPO with channel re-ordering: the flagged channels are the
    ends of every bond, their environment blocks are unit matrices for a canonical random MPS, and <psi|H|psi> through the
    (re-ordered) MPO equals the dense value."""
    import pytdscf_b200 as tb
    from pytdscf_b200._mps_cuda import DeviceMPO, MPSCoefCuda
    from pytdscf_b200.mpo_tools import sop_to_dense, sop_to_mpo

    rng = np.random.default_rng(seed)
    dims = [2, 3, 2, 3, 2]
    n = len(dims)

    def rop(d):
        m = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
        return m + m.conj().T

    terms = [(rng.standard_normal(), {p: rop(dims[p])}) for p in range(n)]
    for _ in range(6):
        a, b = sorted(rng.choice(n, 2, replace=False))
        terms.append((rng.standard_normal(), {int(a): rop(dims[a]), int(b): rop(dims[b])}))
    cores = sop_to_mpo(dims, terms)
    basis = [tb.Exciton(nstate=d) for d in dims]
    model = tb.Model(basis, {"hamiltonian": cores}, bond_dim=4)
    eng = OracleEngine()
    eng.reorder_mpo_channels = True
    H = DeviceMPO(eng, model.hamiltonian)
    flagged = 0
    for terms_p in H.calc_point:
        for t in terms_p:
            assert t.core.l_id in (-1, 0) and t.core.r_id in (-1, t.core.wr - 1)
            flagged += (t.core.l_id >= 0) + (t.core.r_id >= 0)
    assert flagged >= 2 * (n - 2)
    # random MPS, right-canonical around site 0
    from pytdscf_b200._mps_cuda import bond_dims

    tens = []
    for i in range(n):
        ml, mr = bond_dims(dims, i, 4)
        tens.append(eng.to_device(rng.standard_normal((1 if i == 0 else ml, dims[i], 1 if i == n - 1 else mr))
                                  + 1j * rng.standard_normal((1 if i == 0 else ml, dims[i], 1 if i == n - 1 else mr))))
    for i in range(n - 1, 0, -1):
        B, sig = eng.qr_shift("B", tens[i])
        tens[i] = B
        tens[i - 1] = eng.absorb("B", sig, tens[i - 1])
    tens[0] = tens[0] / np.sqrt(eng.inner(tens[0], tens[0], True).real)
    mps = MPSCoefCuda(eng, tens)
    right = mps.construct_op_sites(n - 1, 0, H)
    for p in range(n - 1):
        for t in H.calc_point[p]:
            if t.core.r_id >= 0 and t.key in right[n - 1 - p]:
                E = right[n - 1 - p][t.key].numpy()
                np.testing.assert_allclose(E[:, t.core.r_id, :], np.eye(E.shape[0]), atol=1e-12)
    v = np.asarray(tens[0].numpy())[0]
    for c in tens[1:]:
        v = np.tensordot(v, c.numpy(), axes=(v.ndim - 1, 0))
    v = v.reshape(-1)
    dense = sop_to_dense(dims, terms)
    assert abs(mps.expectation(H) - np.vdot(v, dense @ v)) < 1e-11 * max(1.0, np.abs(dense).max())


def run_adaptive(g, eng, tmp_path, tag):
    import pytdscf_b200 as tb

    os.chdir(tmp_path)
    sim = tb.Simulator(g["name"] + tag, _build_model(g), backend="cuda")
    if eng is not None:
        sim.eng = eng
    sim.set_initial_mps(g["init"])
    Dmax, dD, p_proj, p_svd = g["adaptive"]
    ener, wf = sim.propagate(stepsize=g["dt_au"] * tb.units.au_in_fs, maxstep=g["nstep"], populations=False, record_trace=True,
                             adaptive=True, adaptive_Dmax=int(Dmax), adaptive_dD=int(dD), adaptive_p_proj=p_proj,
                             adaptive_p_svd=p_svd)
    return sim, ener, wf


@pytest.mark.parametrize("case", ["adaptive_exciton", "adaptive_hh6"])
def test_host_adaptive_logic(case, tmp_path):
    """Rank-adaptive one-site TDVP (reference adaptive=True, starting from bond dimension 1 or 2): the product's host
    logic with the oracle's kernels reproduces the reference's bond growth, Krylov trace and observables."""
    g = load_run(case)
    sim, ener, wf = run_adaptive(g, OracleEngine(), tmp_path, "_cpu")
    assert [s.shape for s in wf.ci_coef.sites] == [c.shape for c in g["final"]]
    assert (np.array(wf.ci_coef.trace) == g["trace"]).all()
    for rec, row in zip(sim.history, g["props"], strict=True):
        assert abs(rec["autocorr"] - complex(row[1], row[2])) < 1e-11
        assert abs(rec["energy"] - row[3]) < 1e-12 and abs(rec["norm"] - row[5]) < 1e-12
    from tests.test_gpu_propagation import dense_state

    a, b = dense_state(wf.ci_coef.to_numpy()), dense_state(g["final"])
    assert np.abs(a - b).max() < 1e-10
    # bonddim.dat in the reference's layout, one row per step
    lines = open(os.path.join(g["name"] + "_cpu_prop", "bonddim.dat")).read().splitlines()
    assert lines[0].startswith("# time [fs]") and len(lines) == g["nstep"] + 1
    assert [int(v) for v in lines[-1].split()[1:]] == sim.history[-1]["bonddim"]


@pytest.mark.parametrize("gauge", ["A", "B"])
def test_thin_to_full_matches_scipy_full_qr(gauge):
    """``thin_to_full`` through the padded-matrix QR equals the reference's construction (scipy full QR, first columns
    sign-aligned to the isometry, reference _site_cls.py:294-407) including the ORDER of the complement directions."""
    import scipy.linalg

    from pytdscf_b200._adaptive import thin_to_full
    from pytdscf_b200._mps_cuda import SiteCoef

    rng = np.random.default_rng(5)
    eng = OracleEngine()
    l, c, r, extra = (3, 4, 5, 2) if gauge == "A" else (5, 4, 3, 2)
    if gauge == "A":
        mat = np.linalg.qr(rng.standard_normal((l * c, r)) + 1j * rng.standard_normal((l * c, r)))[0]
        data = mat.reshape(l, c, r)
    else:
        mat = np.linalg.qr(rng.standard_normal((c * r, l)) + 1j * rng.standard_normal((c * r, l)))[0]   # (c r) x l
        data = np.ascontiguousarray(mat.T.reshape(l, c, r))
    full = thin_to_full(eng, SiteCoef(eng.to_device(data), gauge, 0), extra).data.numpy()
    Q, _ = scipy.linalg.qr(mat, mode="full")
    n = mat.shape[1]
    unflip = np.sign(np.sign(np.diag((mat.T.conj() @ Q)[:n, :n]).real) + 0.5)
    Q = Q[:, : n + extra].copy()
    Q[:, :n] *= unflip[None, :]
    ref = Q.reshape(l, c, r + extra) if gauge == "A" else Q.T.reshape(l + extra, c, r)
    np.testing.assert_allclose(full, ref, atol=1e-13)


def test_direct_sum_mpo_is_the_same_operator(tmp_path):
    """``DeviceMPO(merge_terms=True)`` (one direct-sum MPO for all whole-chain keys; the launch-bound small-D option) applies
    the same H_eff: energies / autocorrelation of the golden run to 1e-11, identical Krylov trace, and one term per site."""
    import pytdscf_b200 as tb
    from pytdscf_b200._mps_cuda import DeviceMPO, direct_sum_mpo
    from pytdscf_b200.mpo_tools import mpo_to_dense

    g = load_run("henon_heiles_f6")
    model = _build_model(g)
    H = DeviceMPO(OracleEngine(), model.hamiltonian, merge_terms=True)
    assert H.merged and all(len(t) == 1 for t in H.calc_point)
    # dense check of the direct sum on a small chain
    keys = list(g["operators"].items())
    full = []
    for key, cores in keys:
        full.append([c if c.ndim == 4 else np.einsum("aib,ij->aijb", c, np.eye(c.shape[1])) for c in (np.asarray(x) for x in cores)])
    small = [[c[:, :3, :3, :] for c in row[:3]] for row in full]
    small = [[row[0], row[1], row[2][..., :1] * 0 + row[2].sum(axis=-1, keepdims=True)] for row in small]   # close the chain after 3 sites
    dense_sum = sum(mpo_to_dense(row) for row in small)
    assert np.abs(mpo_to_dense(direct_sum_mpo(small)) - dense_sum).max() < 1e-13
    os.chdir(tmp_path)
    sim = tb.Simulator("merged", model, backend="cuda")
    sim.eng = OracleEngine()
    sim.merge_mpo_terms = True
    sim.set_initial_mps(g["init"])
    ener, wf = sim.propagate(stepsize=g["dt_au"] * tb.units.au_in_fs, maxstep=g["nstep"], thresh_sil=g["thresh_sil"],
                             populations=False, record_trace=True)
    assert (np.array(wf.ci_coef.trace) == g["trace"]).all()
    for rec, row in zip(sim.history, g["props"], strict=True):
        assert abs(rec["autocorr"] - complex(row[1], row[2])) < 1e-11
        assert abs(rec["energy"] - row[3]) < 1e-11 * max(1.0, abs(row[3]))


def test_single_site_chain_reproduces_the_reference_literal(tmp_path):
    """The f = 1 row of the reference's tests/test_henon_heiles.py:21 (omega = 4000 cm-1, N = 5 HO-DVR points, m = 4, dt = 0.01 fs,
    first excited state, pinned energy 0.027338011517478895 = 1.5 omega) -- a chain of ONE site: no bond, no QR shift, no K step.
    The potential of one mode needs no MPO builder (a (1, N, 1) diagonal core on the grid), the kinetic MPO is this package's."""
    import pytdscf_b200 as tb

    os.chdir(tmp_path)
    w, N = 4000, 5
    prim = [tb.HarmonicOscillator(N, w)]
    q = np.array(prim[0].get_grids())
    V = [((w / tb.units.au_in_cm1) ** 2 / 2 * q**2).reshape(1, N, 1)]
    model = tb.Model(prim, operators={"potential": V, "kinetic": tb.construct_kinetic_mpo(prim)}, bond_dim=4)
    model.init_weight_VIBSTATE = [[[0.0, 1.0] + [0.0] * (N - 2)]]
    sim = tb.Simulator(jobname="henon_heiles", model=model, backend="cuda", verbose=0)
    sim.eng = OracleEngine()
    ener, wf = sim.propagate(maxstep=3, stepsize=0.01)
    assert ener == pytest.approx(0.027338011517478895)            # the reference's own tolerance (rel 1e-6); measured 7e-16
    assert [tuple(s.data.shape) for s in wf.ci_coef.sites] == [(1, N, 1)]
    assert all(abs(abs(rec["autocorr"]) - 1.0) < 1e-12 for rec in sim.history)      # an eigenstate only picks up a phase


@pytest.mark.parametrize("improved", [True, False])
def test_harmonic_relaxation_reproduces_the_reference_literal(improved, tmp_path):
    """The reference's tests/test_harmonic_dvr_func_full_mpssm_jax.py:17-56 (three HO-DVR modes of 1500 / 2000 / 2500 cm-1 with
    five points, harmonic potential, m_aux_max = 4, ``relax(maxstep=3, stepsize=0.1)``) pins the zero-point energy
    0.013669005758739458.  The separable potential needs no MPO builder: three one-site diagonal keys plus this package's kinetic
    MPO.  ``improved=True`` is the reference's default (per-site Lanczos eigen-solver), ``False`` imaginary time."""
    import pytdscf_b200 as tb

    os.chdir(tmp_path)
    freqs = [1500, 2000, 2500]
    prim = [tb.HarmonicOscillator(5, w, 0.0) for w in freqs]
    pot = {}
    for i, (p, w) in enumerate(zip(prim, freqs)):
        q = np.array(p.get_grids())
        pot[(i,)] = tb.TensorOperator(mpo=[((w / tb.units.au_in_cm1) ** 2 / 2 * q**2).reshape(1, 5, 1)], legs=(i,))
    kin = {tuple((i, i) for i in range(3)): tb.TensorOperator(mpo=tb.construct_kinetic_mpo(prim))}
    ham = tb.TensorHamiltonian(ndof=3, potential=[[pot]], kinetic=[[kin]], backend="cuda")
    model = tb.Model(tb.BasInfo([prim]), {"hamiltonian": ham})
    model.m_aux_max = 4
    sim = tb.Simulator("harmonic_dvr", model, backend="cuda", verbose=0)
    sim.eng = OracleEngine()
    ener, wf = sim.relax(maxstep=3, stepsize=0.1, improved=improved)
    assert ener == pytest.approx(0.013669005758739458)            # rel 1e-6 in the reference's test; measured 7e-16 / 2e-16


@pytest.mark.parametrize("name", ["exciton_D6", "h2co_D16", "liouville_spin3", "gate_exciton_D6", "adaptive_hh6"])
def test_host_logic_with_the_device_conventions(name, tmp_path):
    """The serial goldens once more with the engine following the DEVICE's conventions where they differ from the oracle's
    (tests/device_numerics_engine.py: MPO channels reordered so that identity channels come first / last, which also changes the
    order of the term sums; SVD-based calls by one-sided Jacobi): identical Krylov traces and bond growth, observables within
    the bars of the GPU parity tests."""
    from tests.device_numerics_engine import DeviceNumericsEngine

    g = load_run(name)
    if g["adaptive"] is not None:
        sim, ener, wf = run_adaptive(g, DeviceNumericsEngine(), tmp_path, "_dev")
    else:
        import pytdscf_b200 as tb

        os.chdir(tmp_path)
        sim = tb.Simulator(name + "_dev", _build_model(g), backend="cuda", verbose=0)
        sim.eng = DeviceNumericsEngine()
        sim.set_initial_mps(g["init"])
        hil = g["space"] == "hilbert"
        ener, wf = sim.propagate(stepsize=g["dt_au"] * tb.units.au_in_fs, maxstep=g["nstep"], thresh_sil=g["thresh_sil"],
                                 integrator=g["integrator"], conserve_norm=g["conserve_norm"], energy=hil, autocorr=hil,
                                 norm=hil, populations=hil, record_trace=True)
    from tests.test_gpu_propagation import dense_state, tolerances

    # the GPU tests' own bars: 1e-10, or 4 x the reference algorithm's rounding-noise floor where that is larger (exciton_D6 and
    # its gate variant: bond dimension 6 on a state of numerical rank ~2, tests/golden/noise_floor.json)
    tol = tolerances(name if name in ("exciton_D6", "h2co_D16", "liouville_spin3") else "exciton_D6" if "exciton_D6" in name
                     else "henon_heiles_f6")
    assert (np.array(wf.ci_coef.trace) == g["trace"]).all()
    assert [s.shape for s in wf.ci_coef.sites] == [c.shape for c in g["final"]]
    if g["space"] == "hilbert":
        for rec, row in zip(sim.history, g["props"], strict=True):
            assert abs(rec["autocorr"] - complex(row[1], row[2])) < tol["autocorr"]
            assert abs(rec["energy"] - row[3]) < tol["energy"] * max(1.0, abs(row[3]))
            assert abs(rec["norm"] - row[5]) < tol["autocorr"]
    a, b = dense_state(wf.ci_coef.to_numpy()), dense_state(g["final"])
    assert np.abs(a - b).max() < tol["state"] * max(1.0, np.abs(b).max())
