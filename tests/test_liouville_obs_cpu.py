"""CPU host-logic test of the Liouville-space expectation values Tr(O rho) (reference ``_exp_liouville``,
pytdscf/_mps_cls.py:3769-3838) and of the sub-space projection of the initial MPDO (``project_subspace``,
pytdscf/_mps_mpo.py:196-220) against goldens of the unmodified reference (tests/golden/liouville_obs.npz), with the
oracle's NumPy kernels injected in place of the CUDA engine."""
import pytest

from oracle.oracle_engine import OracleEngine
from tests.liouville_obs_cases import check_case, run_case


@pytest.mark.parametrize("tag", ["full", "sub"])
def test_liouville_observables_and_subspace_host_logic(tag, tmp_path):
    sim, wf = run_case(tag, tmp_path, engine=OracleEngine())
    check_case(tag, sim, wf, tol_expect=1e-13, tol_state=1e-12)
