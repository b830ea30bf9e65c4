"""Primitive bases -- only what the TDVP hot path and its input preparation need.

The hot path uses ``len(basis)`` (the physical dimension d of a site) and, for HO-DVR sites without an
explicit Hartree product, the FBR->DVR unitary (reference: ``pytdscf/_mps_mpo.py:96-110``).  The d x d
operator matrices are used to BUILD MPO cores (host-side input preparation).
Reference: ``pytdscf/basis/{exciton,boson,ho,abc}.py``.
"""
from __future__ import annotations

import math

import numpy as np

from . import units


class Exciton:
    """n-level site (electronic states, spins, vectorised density-matrix indices)."""

    def __init__(self, nstate: int, names: list[str] | None = None):
        self.nstate = int(nstate)
        self.names = [f"S{i}" for i in range(nstate)] if names is None else list(names)
        if len(self.names) != nstate:
            raise ValueError("len(names) != nstate")

    def get_annihilation_matrix(self) -> np.ndarray:
        return np.diag(np.ones(self.nstate - 1), k=1)

    def get_creation_matrix(self) -> np.ndarray:
        return self.get_annihilation_matrix().T

    @property
    def nprim(self) -> int:
        return self.nstate

    def __len__(self) -> int:
        return self.nstate


class Boson:
    """Truncated harmonic mode in the number basis."""

    def __init__(self, nstate: int):
        self.nstate = int(nstate)

    def get_annihilation_matrix(self) -> np.ndarray:
        return np.diag(np.sqrt(np.arange(1, self.nstate, dtype=float)), k=1)

    def get_creation_matrix(self) -> np.ndarray:
        return self.get_annihilation_matrix().T

    def get_number_matrix(self) -> np.ndarray:
        return np.diag(np.arange(self.nstate, dtype=float))

    def get_q_matrix(self) -> np.ndarray:
        a = self.get_annihilation_matrix()
        return (a + a.T) / math.sqrt(2.0)

    def get_p_matrix(self) -> np.ndarray:
        a = self.get_annihilation_matrix()
        return 1.0j * (a.T - a) / math.sqrt(2.0)

    def _ladder_sum_squared(self, sign: float) -> np.ndarray:
        """(b+ + sign b)^2 evaluated in a space one level larger and cut back, so that the last diagonal element is the exact
        matrix element instead of the truncation artefact of squaring the cut matrices (reference ``margin=1``,
        pytdscf/basis/boson.py:129-153)."""
        b = np.diag(np.sqrt(np.arange(1, self.nstate + 1, dtype=float)), k=1)
        s = b.T + sign * b
        return (s @ s)[:-1, :-1]

    def get_q2_matrix(self) -> np.ndarray:
        """q^2 = (b+ + b)^2 / 2."""
        return 0.5 * self._ladder_sum_squared(+1.0)

    def get_p2_matrix(self) -> np.ndarray:
        """p^2 = -(b+ - b)^2 / 2."""
        return -0.5 * self._ladder_sum_squared(-1.0)

    @property
    def nprim(self) -> int:
        return self.nstate

    def __len__(self) -> int:
        return self.nstate


class HarmonicOscillator:
    """Harmonic-oscillator DVR in mass-weighted coordinates (grid = eigenvalues of the position matrix in the
    HO eigenbasis; MCTDH review, Phys. Rep. 324, 1 (2000), appendix B)."""

    def __init__(self, ngrid: int, omega: float, q_eq: float = 0.0, units_: str = "cm-1", **kwargs):
        units_ = kwargs.pop("units", units_)
        if kwargs:
            raise TypeError(f"unexpected arguments {list(kwargs)}")
        self.ngrid = self.nprim = int(ngrid)
        u = units_.lower()
        if u in ("cm1", "cm-1", "kaiser"):
            self.omega = omega / units.au_in_cm1
        elif u in ("au", "hartree", "a.u."):
            self.omega = omega
        elif u == "ev":
            self.omega = omega / units.au_in_eV
        else:
            raise ValueError(f"{units_} must be one of cm-1, au, eV")
        self.q_eq = q_eq
        self._grids = None
        self._unitary = None

    def __len__(self) -> int:
        return self.ngrid

    def get_pos_rep_matrix(self) -> np.ndarray:
        n = self.ngrid
        off = np.sqrt(np.arange(1, n) / 2.0 / self.omega).astype(complex)
        return np.diag(np.full(n, self.q_eq, dtype=complex)) + np.diag(off, 1) + np.diag(off, -1)

    def _diagonalise(self):
        if self._grids is None:
            from scipy.linalg import eigh

            val, vec = eigh(self.get_pos_rep_matrix())
            # sign convention: positive quadrature weights, sqrt(w_a) = conj(U[0,a]) / phi_0(x_a)  (phi_0 > 0)
            for a in range(self.ngrid):
                if np.conjugate(vec[0, a]).real < 0:
                    vec[:, a] *= -1.0
            self._grids, self._unitary = list(val), vec
        return self._grids, self._unitary

    def get_grids(self) -> list[float]:
        return self._diagonalise()[0]

    def get_unitary(self) -> np.ndarray:
        """u[j, alpha]: FBR index j (HO quantum number) x DVR index alpha."""
        return self._diagonalise()[1]

    def get_1st_derivative_matrix_fbr(self) -> np.ndarray:
        """First-derivative matrix in the HO eigenbasis WITH THE REFERENCE'S SIGN (pytdscf/basis/ho.py:130-145):
        D[j, j+1] = -sqrt(omega (j+1) / 2), D[j+1, j] = +sqrt(omega (j+1) / 2).  With the position matrix of
        ``get_pos_rep_matrix`` this D satisfies [D, q] = -1, i.e. it is -d/dq for real HO functions; MPO cores built from it
        must equal the reference's, so its convention is kept (the second-derivative matrix does not depend on it)."""
        up = np.sqrt(self.omega * np.arange(1, self.ngrid) / 2.0)
        return np.diag(up, -1) - np.diag(up, 1)

    def get_1st_derivative_matrix_dvr(self) -> np.ndarray:
        U = self.get_unitary()
        return U.conj().T @ self.get_1st_derivative_matrix_fbr() @ U

    def get_2nd_derivative_matrix_fbr(self) -> np.ndarray:
        n = self.ngrid
        diag = -self.omega / 2 * (2 * np.arange(n) + 1)
        off2 = self.omega / 2 * np.sqrt(np.arange(1, n - 1) * np.arange(2, n))
        return np.diag(diag) + np.diag(off2, 2) + np.diag(off2, -2)

    def get_2nd_derivative_matrix_dvr(self) -> np.ndarray:
        U = self.get_unitary()
        return U.conj().T @ self.get_2nd_derivative_matrix_fbr() @ U
