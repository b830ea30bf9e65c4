"""Unit conversion factors (CODATA via scipy.constants), same names as ``pytdscf/units.py``:
``xxx_in_yyy = z`` means 1 [xxx] = z [yyy]."""
from scipy.constants import physical_constants as _pc

au_in_cm1 = _pc["atomic unit of energy"][0] / (_pc["speed of light in vacuum"][0] * 1.0e02) / _pc["Planck constant"][0]
au_in_fs = _pc["atomic unit of time"][0] / 1.0e-15
au_in_eV = _pc["Hartree energy in eV"][0]
au_in_dalton = _pc["electron mass"][0] / _pc["atomic mass constant"][0]
au_in_angstrom = _pc["Bohr radius"][0] / 1.0e-10
au_in_debye = _pc["atomic unit of electric dipole mom."][0] * _pc["speed of light in vacuum"][0] * 1.0e21
# aliases the reference exports as well
Hartree_in_cm1 = au_in_cm1
Has_in_eV = au_in_eV
au_in_AMU = au_in_dalton
Bohr_in_angstrom = au_in_angstrom
