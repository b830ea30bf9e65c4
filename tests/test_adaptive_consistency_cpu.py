"""Consistency of the rank-adaptive building blocks, written after the reference's own tests/test_a1tdvp.py:176-329 (its hand
-rolled adaptive sweep): at every site of a forward and a backward sweep
  * the new bond dimension grows by at most ``dD`` (``:247-253``),
  * propagating the site INTO the enlarged tensor (``tensor_shapes_out``) agrees with the un-enlarged propagation on the common
    block to 1e-3 (``:288-291``),
  * the bond matrix of the following QR is newD x newD (``:301-303``) and the shifted site is an isometry (``:325-326``).
Host logic of ``pytdscf_b200/_adaptive.py`` + ``MPSCoefCuda`` with the oracle's kernels injected, on the exciton model of that
test (bond dimension 1 at the start, Dmax 100, dD 30, p_proj 1e-5)."""
import numpy as np
import pytest

from oracle.oracle_engine import OracleEngine
from pytdscf_b200 import _adaptive as ad
from pytdscf_b200._const_cls import RunConfig
from pytdscf_b200._mps_cuda import DeviceMPO, MPSCoefCuda, SiteCoef
from tests.golden_io import load_run
from tests.test_host_sweep_cpu import _build_model


@pytest.mark.parametrize("dD", [30, 2])
def test_adaptive_sweep_consistency(dD):
    g = load_run("adaptive_exciton")
    eng = OracleEngine()
    model = _build_model(g)
    H = DeviceMPO(eng, model.hamiltonian)
    mps = MPSCoefCuda(eng, [eng.to_device(c) for c in g["init"]])
    cfg = RunConfig(adaptive=True, Dmax=100, dD=dD, p_proj=1.0e-05, p_svd=1.0e-07)
    dt = g["dt_au"]
    n = mps.nsite
    grew = False
    for to in ("->", "<-"):
        fwd = to == "->"
        begin, end, step = (0, n - 1, 1) if fwd else (n - 1, 0, -1)
        sites = mps.sites
        full = ad.get_superblock_full(eng, sites, cfg.dD)
        op_sys = mps.construct_op_zerosite()
        env_sites = mps.construct_op_sites(end, begin, H) if mps.op_sys_sites is None else mps.op_sys_sites[:]
        mps.op_sys_sites = [op_sys]
        for p in range(begin, end + step, step):
            op_env = env_sites.pop()
            l, c, r = sites[p].shape  # noqa: E741
            if p != end:
                newD, _err, op_env_bra, op_env_braket = ad.get_adaptive_rank_and_block(mps, p, full, env_sites[-1], H, to, cfg)
                assert newD <= (r if fwd else l) + cfg.dD
                grew = grew or newD > (r if fwd else l)
            else:
                newD, op_env_bra, op_env_braket = 1, op_env, op_env
            shape_out = (l, c, newD) if fwd else (newD, c, r)
            x = sites[p].data
            plain = mps._expm(cfg, -1.0j, dt, x, p, 0, hterms=mps.operators_for_superH(p, op_sys, op_env, H, fwd))
            grown = mps._expm(cfg, -1.0j, dt, x, p, 0, shape_out=shape_out, hterms=mps.operators_for_superH(p, op_sys, op_env_bra, H, fwd))
            assert tuple(grown.shape) == shape_out
            np.testing.assert_allclose(np.asarray(grown)[:l, :c, :r], np.asarray(plain), atol=1e-03)
            sites[p] = SiteCoef(grown, "Psi", p)
            if p == end:
                break
            gauge = "A" if fwd else "B"
            iso, sigma = eng.qr_shift(gauge, grown)
            assert tuple(sigma.shape) == (newD, newD)
            m = np.asarray(iso).reshape(-1, newD) if fwd else np.asarray(iso).reshape(newD, -1).T
            np.testing.assert_allclose(m.conj().T @ m, np.eye(newD), atol=1e-12)
            sites[p] = SiteCoef(iso, gauge, p)
            op_sys = mps.renormalize_op_psite(p, op_sys, H, fwd)
            sigma = mps._expm(cfg, +1.0j, dt, sigma, p, 1, kterms=mps.operators_for_superK(op_sys, op_env_braket, H, fwd))
            q = p + step
            sites[q] = SiteCoef(eng.absorb(gauge, sigma, sites[q].data), "Psi", q)
            mps.op_sys_sites.append(op_sys)
        # the state stays normalised through the sweep (unitary local steps on a growing manifold)
        psi = np.asarray(sites[0].data)
        for s in sites[1:]:
            psi = np.tensordot(psi, np.asarray(s.data), axes=(-1, 0))
        assert abs(np.linalg.norm(psi) - 1.0) < 1e-6
    assert grew, "the bonds of the product state must grow under the exciton Hamiltonian"
