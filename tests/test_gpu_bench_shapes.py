"""GPU parity at the shapes bench.py measures: every GEMM configuration FORCED through the ABI knob
``tdvp_set_gemm_config`` on the config-3/4/5 GEMM shapes, a ``CheckedEngine`` replay (tests/checked_engine.py: every
engine call re-run on the oracle's NumPy kernels with the same inputs) of one full site update at the config-3, config-4
and config-5 site shapes, and one full time step of config 2 against ``TDVPOracle``.

Reference functions covered at these sizes: ``_op_lcr_dot`` / ``_op_lr_dot`` (pytdscf/_contraction.py:1038-1176, 1297-1352),
``contract_with_site_mpo`` (:148-397), ``short_iterative_lanczos`` / ``_arnoldi`` (pytdscf/_integrator.py:287-655),
``gauge_trf`` (pytdscf/_site_cls.py:138-292), ``trans_next_psite_APsiB`` (pytdscf/_mps_cls.py:1172-1206)."""
import numpy as np
import pytest
import torch

from oracle import tdvp_oracle as orc
from pytdscf_b200 import workloads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from pytdscf_b200._engine import Engine

    e = Engine(0)
    yield e
    e.set_gemm_config("auto", 0, 0)
    e.close()


# (label, M, N, K, transA, transB): the GEMMs of one H_eff / K_eff / env update at the benchmarked sites
# (scripts/microbench/zgemm_shapes.py); stage 2 (two-level rows) is covered by the site-update replays below
GEMM_SHAPES = [
    ("c4 heff stage1 NN", 8192, 4096, 1024, 0, 0),
    ("c4 heff stage3 NT", 4096, 1024, 8192, 0, 1),
    ("c4 env step3 CN", 1024, 8192, 4096, 2, 0),
    ("c4 keff gemm1 NN", 8192, 1024, 1024, 0, 0),
    ("c4 keff gemm2 NT", 1024, 1024, 8192, 0, 1),
    ("c4 id-skip stage1 (7 channels)", 7168, 4096, 1024, 0, 0),
    ("c3 heff stage1 NN", 1280, 2560, 256, 0, 0),
    ("c3 heff stage3 NT", 2560, 256, 1280, 0, 1),
    ("c3 env step3 CN", 256, 1280, 2560, 2, 0),
    ("c3 keff gemm1 NN", 1280, 256, 256, 0, 0),
    ("c3 keff gemm2 NT", 256, 256, 1280, 0, 1),
    ("c5 heff stage1 NN pot", 1536, 4096, 512, 0, 0),
    ("c5 heff stage3 NT pot", 4096, 512, 1536, 0, 1),
    ("c5 keff gemm2 NT", 512, 512, 1536, 0, 1),
    ("qr VhC (skinny, split-K)", 1024, 32, 4096, 1, 0),
    ("c2 heff stage1 NN", 192, 512, 64, 0, 0),
    ("c2 heff stage3 NT (cluster split-K)", 512, 64, 192, 0, 1),
    ("c2 env step3 CN", 64, 192, 500, 2, 0),
    ("stream-K: 150 tiles on 148 SMs", 1280, 960, 512, 0, 0),
    ("stream-K: ragged M / N / K tails", 2500, 1100, 1000, 0, 1),
]
# (tile, splitk, c_stream) -- see tdvp_set_gemm_config
# split-K of the tiny / small tiles runs as a thread-block cluster of S <= 16 CTAs (partials reduced through distributed
# shared memory); c_stream = 1 (or the big tile) keeps the scratch + reduction-kernel path under test
GEMM_CONFIGS = [("auto", 0, 0), ("big", 1, 2), ("big", 1, 1), ("small", 1, 2), ("tiny", 1, 2), ("big", 4, 0), ("tiny", 3, 0),
                ("tiny", 8, 0), ("small", 4, 0), ("tiny", 3, 1), ("tiny", 16, 0), ("tiny", 12, 1),
                ("tma", 1, 0), ("tma", 2, 0), ("tma_tiles", 1, 0)]


@pytest.mark.parametrize("cfg", GEMM_CONFIGS, ids=lambda c: f"{c[0]}-S{c[1]}-cs{c[2]}")
@pytest.mark.parametrize("shape", GEMM_SHAPES, ids=lambda s: s[0].replace(" ", "_"))
def test_zgemm_forced_config_on_bench_shapes(eng, shape, cfg):
    """C = alpha op(A) op(B) + beta C against an fp64 reference (torch.matmul complex128 = cuBLAS ZGEMM, used only as the
    checker), relative error <= 1e-13 of max|C| for every forced configuration."""
    label, M, N, K, ta, tb = shape
    g = torch.Generator(device="cuda").manual_seed(M + 3 * N + 7 * K + ta + 2 * tb)

    def crand(*s):
        return torch.view_as_complex(torch.randn(*s, 2, dtype=torch.float64, device="cuda", generator=g)) / np.sqrt(2)

    A = crand(*((K, M) if ta else (M, K)))
    B = crand(*((N, K) if tb else (K, N)))
    C0 = crand(M, N)
    alpha, beta = 0.7 - 0.2j, -0.3 + 0.5j
    opA = {0: A, 1: A.T, 2: A.conj().T}[ta]
    opB = {0: B, 1: B.T, 2: B.conj().T}[tb]
    ref = alpha * torch.matmul(opA, opB) + beta * C0
    C = C0.clone()
    eng.set_gemm_config(*cfg)
    try:
        eng.zgemm(A, B, ta, tb, alpha, beta, C)
    finally:
        eng.set_gemm_config("auto", 0, 0)
    err = float((C - ref).abs().max() / ref.abs().max())
    assert err <= 1e-13, (label, cfg, err)


# -------------------------------------------------------------------------------------------------------------
def _herm_block(rng, D, w, scale):
    x = (rng.standard_normal((D, w, D)) + 1j * rng.standard_normal((D, w, D))) * scale
    return (x + x.conj().transpose(2, 1, 0)) / 2


def _site_of(wl, want_d):
    """A full-bond-dimension site of physical dimension ``want_d`` (Dl = Dr = D)."""
    from pytdscf_b200._mps_cuda import bond_dims

    for i, d in enumerate(wl.dims):
        if d == want_d and bond_dims(wl.dims, i, wl.bond_dim) == (wl.bond_dim, wl.bond_dim) and 0 < i < len(wl.dims) - 1:
            return i
    raise AssertionError("no such site")


SITE_CASES = [
    # workload, physical dimension of the sampled site, GEMM configuration forced during the replay
    ("c3", 10, ("auto", 0, 0)),
    ("c5", 8, ("auto", 0, 0)),
    ("c4", 4, ("auto", 0, 0)),
    ("c4", 4, ("tma", 0, 0)),
    ("c3", 10, ("tma", 0, 0)),
]


@pytest.mark.parametrize("name,d,cfg", SITE_CASES, ids=lambda v: "-".join(map(str, v)) if isinstance(v, tuple) else str(v))
def test_site_update_replay_at_bench_shape(eng, name, d, cfg):
    """One site update of the sweep -- H solve, QR shift, environment update of every MPO term, K solve, absorb -- at a
    full-D site of the named BASELINE configuration (c3: D = 256, d = 10; c5: D = 512, d = 8; c4: D = 1024, d = 4,
    Arnoldi), on dense seeded blocks with the workload's own MPO cores.  Every engine call is replayed on the oracle's
    NumPy kernels from the same inputs: per-kernel deviation <= 1e-11 and identical Krylov iteration counts."""
    from tests.checked_engine import CheckedEngine

    wl = workloads.by_name(name)
    p = _site_of(wl, d)
    D = wl.bond_dim
    H = orc.MPOHamiltonian(len(wl.dims), wl.operators, wl.coupleJ)
    rng = np.random.default_rng(20 + len(name) + d)
    ce = CheckedEngine(eng)
    eng.set_gemm_config(*cfg)
    try:
        psi = rng.standard_normal((D, d, D)) + 1j * rng.standard_normal((D, d, D))
        psi /= np.linalg.norm(psi)
        hterms = []
        left_blocks = []
        for core in H.calc_point[p]:
            w_l, w_r = core.data.shape[0], core.data.shape[-1]
            # block scale: ||L x W x R|| of order one so that the Krylov solves converge in 6-12 vectors at |scale| = 1
            L = None if core.is_left else eng.to_device(_herm_block(rng, D, w_l, 0.5 / np.sqrt(D * w_l)))
            R = None if core.is_right else eng.to_device(_herm_block(rng, D, w_r, 0.5 / np.sqrt(D * w_r)))
            dc = eng.upload_core(core.data)
            hterms.append((L, dc, R, 1.0))
            left_blocks.append((L, dc, R))
        x = eng.to_device(psi)
        # scale of the local generator from two power iterations (GPU only; the checked solve below is what is compared)
        y = eng.heff_apply(hterms, x)
        nrm = float(torch.linalg.vector_norm(eng.heff_apply(hterms, y / torch.linalg.vector_norm(y))))
        scale = -1.0j * 0.7 / nrm
        kind = wl.integrator
        n_h = ce.krylov_expm(kind, scale, 1e-9, 0, wl.conserve_norm, x, hterms=hterms)
        # warm-up gated path (skipped at D = 1024, where every oracle solve costs ~30 s of host time)
        n_h2 = ce.krylov_expm(kind, scale, 1e-9, max(0, n_h - 2), wl.conserve_norm, x, hterms=hterms) if D <= 512 else -1
        A, sigma = ce.qr_shift("A", x)
        kterms = []
        for (L, dc, R) in left_blocks:
            E = ce.env_update("A", A, A, L, dc)
            if R is not None:
                kterms.append((E, R, 1.0))
        y = eng.keff_apply(kterms, sigma)
        nrm_k = float(torch.linalg.vector_norm(y) / torch.linalg.vector_norm(sigma))
        n_k = ce.krylov_expm(kind, 1.0j * 0.7 / max(nrm_k, 1e-300), 1e-9, 0, wl.conserve_norm, sigma, kterms=kterms)
        nxt = eng.to_device(rng.standard_normal((D, d, D)) + 1j * rng.standard_normal((D, d, D)))
        nxt = ce.absorb("A", sigma, nxt)
        if D <= 512:
            ce.heff_apply(hterms, nxt)
            Bs, sig_b = ce.qr_shift("B", nxt)
            for (L, dc, R) in left_blocks[:1]:
                if R is not None:
                    ce.env_update("B", Bs, Bs, R, dc)
    finally:
        eng.set_gemm_config("auto", 0, 0)
    print(f"\n{name} site {p} (D={D}, d={d}) cfg={cfg}: n_H={n_h}, {n_h2}, n_K={n_k}\n" + ce.report())
    assert 3 <= n_h <= 20 and 2 <= n_k <= 20
    assert ce.dev["krylov_niter"] == 0, ce.worst["krylov_niter"]
    for key in ("env_update", "krylov_expm", "absorb", "qr_shift_product") + (("heff_apply",) if D <= 512 else ()):
        assert ce.dev[key] <= 1e-11, (key, ce.dev[key], ce.worst[key])
    assert ce.dev["qr_shift_isometry"] <= 1e-12


# -------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("merge", [False, True], ids=["per-key", "direct-sum"])
def test_c2_full_step_matches_oracle(eng, merge):
    """One full time step (2 half sweeps, 64 sites, D = 64, 254 Krylov solves) of BASELINE config 2 -- the launch-bound
    regime that uses the tiny GEMM tiles, cluster split-K and the fused Lanczos step -- against ``TDVPOracle`` on the host.

    The oracle's own Krylov trace is not stable under one-ulp perturbations at this size (tests/golden/noise_floor_c2.json:
    2 of 3 seeded 2e-16 perturbations of the H_eff outputs move one stop decision by +-1 and the autocorrelation by
    5e-12), so the trace may differ from the oracle's in at most 3 solves by at most one vector; energy and
    autocorrelation must agree to 1e-10.  ``merge``: the opt-in direct-sum MPO (one GEMM chain for potential + kinetic
    terms, ``DeviceMPO(merge_terms=True)``), which bench.py uses for the D <= 64 workloads."""
    from pytdscf_b200._const_cls import RunConfig
    from pytdscf_b200._mps_cuda import DeviceMPO, MPSCoefCuda

    wl = workloads.by_name("c2")
    Ho = orc.MPOHamiltonian(len(wl.dims), wl.operators, wl.coupleJ)
    o = orc.TDVPOracle(Ho, orc.initial_mps(wl.dims, wl.bond_dim, wl.hartree, space=wl.space), integrator=wl.integrator,
                       conserve_norm=wl.conserve_norm, space=wl.space)
    model = wl.model()
    H = DeviceMPO(eng, model.hamiltonian, merge_terms=merge)
    assert H.merged == merge
    mps = MPSCoefCuda.alloc_random(eng, model)
    for c, r in zip(mps.to_numpy(), o.mps, strict=True):
        np.testing.assert_allclose(c, r, rtol=0, atol=1e-13)
    cfg = RunConfig(jobname="c2", space=wl.space, integrator=wl.integrator, conserve_norm=wl.conserve_norm)
    mps.record_trace = True
    e0, a0 = o.expectation().real, o.autocorr()
    assert abs(mps.expectation(H).real - e0) <= 1e-10 * abs(e0)
    mps.propagate(wl.dt_au, H, cfg)
    o.propagate(wl.dt_au)
    got = [tuple(t) for t in mps.trace]
    ref = [(0 if k == "H" else 1, s, n) for k, s, n in o.trace]
    assert len(got) == len(ref) and all(g[:2] == r[:2] for g, r in zip(got, ref))
    diff = [(g, r) for g, r in zip(got, ref) if g[2] != r[2]]
    print(f"\nc2 step: {len(ref)} solves, {len(diff)} Krylov counts differ from the oracle: {diff}")
    assert len(diff) <= 3 and all(abs(g[2] - r[2]) <= 1 for g, r in diff), diff
    e1, a1 = o.expectation().real, o.autocorr()
    assert abs(mps.expectation(H).real - e1) <= 1e-10 * abs(e1)
    assert abs(mps.autocorr() - a1) <= 1e-10
    assert abs(mps.norm() - o.norm()) <= 1e-10
