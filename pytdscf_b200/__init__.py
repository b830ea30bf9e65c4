"""B200-native TDVP backend (``backend="cuda"``) behind a PyTDSCF-style Model / Simulator API.

Only the MPS/MPO one-site TDVP propagation hot path of PyTDSCF is implemented; it runs
entirely on the GPU through ``libtdvp_b200.so`` (``include/tdvp_b200.h``).  There is no CPU fallback.
"""
from . import kraus, units, util
from .basis import Boson, Exciton, HarmonicOscillator
from .dvr_operator_cls import TensorOperator, construct_kinetic_mpo, construct_kinetic_operator
from .hamiltonian_cls import TensorHamiltonian
from .model_cls import BasInfo, Model
from .checkpoint import export_mpo_npz, import_mpo_npz, read_reference_wavefunction, write_reference_wavefunction
from .simulator_cls import Simulator

__version__ = "0.1.0"
__all__ = ["kraus", "util", "export_mpo_npz", "import_mpo_npz", "read_reference_wavefunction", "write_reference_wavefunction", "units", "Boson", "Exciton", "HarmonicOscillator", "TensorOperator", "construct_kinetic_mpo", "construct_kinetic_operator",
           "TensorHamiltonian", "BasInfo", "Model", "Simulator", "__version__"]
