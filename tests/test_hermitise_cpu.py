"""CPU host-logic test of ``MPSCoefCuda.hermitise`` (reference ``MPSCoef.hermitise`` / ``svd_conj_mpdo``,
pytdscf/_mps_cls.py:2289-2312, 2516-2562) against the unmodified reference's output, with the oracle's NumPy kernels
injected in place of the CUDA engine."""
import numpy as np
import pytest

from oracle.oracle_engine import OracleEngine
from tests.device_numerics_engine import DeviceNumericsEngine
from tests.hermitise_cases import run_and_check


@pytest.mark.parametrize("engine", [OracleEngine, DeviceNumericsEngine])   # LAPACK's SVD conventions / the device's
@pytest.mark.parametrize("tag", ["prop", "rand"])
def test_hermitise_host_logic(tag, engine):
    err, asym = run_and_check(tag, engine(), tol=1e-12)
    if tag == "prop":
        assert asym < 1e-12     # full bonds: (rho + rho^dagger) / 2 is kept exactly


def test_hermitise_single_site_and_subspace_guard():
    from pytdscf_b200._mps_cuda import MPSCoefCuda

    eng = OracleEngine()
    rng = np.random.default_rng(5)
    a = rng.standard_normal((1, 9, 1)) + 1j * rng.standard_normal((1, 9, 1))
    mps = MPSCoefCuda(eng, [eng.to_device(a)])
    mps.hermitise()
    rho = np.asarray(mps.sites[0].data).reshape(3, 3)
    assert np.allclose(rho, 0.5 * (a.reshape(3, 3) + a.reshape(3, 3).conj().T), atol=1e-15)
    mps.subspace = {0: (9, (0, 4, 8))}
    with pytest.raises(NotImplementedError):
        mps.hermitise()
