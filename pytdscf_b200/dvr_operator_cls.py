"""MPO containers handed to ``TensorHamiltonian`` (reference: ``pytdscf/dvr_operator_cls.py:40-160, 1199-1250``).

Only the ``mpo=[...]`` form is supported: building / compressing MPOs from grid potentials is one-off CPU
preprocessing in the reference and out of the hot-path scope; the hot path consumes
``list[np.ndarray]`` cores, 3-index (w_l, d, w_r) = diagonal, 4-index (w_l, d_bra, d_ket, w_r) = full.
"""
from __future__ import annotations

import numpy as np


class TensorOperator:
    def __init__(self, *, mpo: list[np.ndarray] | None = None, legs: tuple[int, ...] | None = None,
                 name: str | None = None, **unsupported):
        if mpo is None:
            raise NotImplementedError(
                "TensorOperator(tensor=...) needs the reference's MPO compressor; pass decomposed cores with mpo=[...]")
        if unsupported:
            raise TypeError(f"unsupported arguments: {sorted(unsupported)}")
        if not isinstance(mpo, list) or not mpo:
            raise TypeError("mpo must be a non-empty list of arrays")
        for core in mpo:
            if np.ndim(core) not in (3, 4):
                raise ValueError(f"core.ndim must be 3 or 4, but {np.ndim(core)}")
        self.tensor_decomposed = [np.asarray(c) for c in mpo]
        self.only_diag = all(c.ndim == 3 for c in self.tensor_decomposed)
        self.bond_dimension = [1] + [c.shape[-1] for c in self.tensor_decomposed] + [1]
        self.shape = tuple(i for c in self.tensor_decomposed for i in c.shape[1:-1])
        self.name = name
        if legs is None:
            _legs: list[int] = []
            for i, c in enumerate(self.tensor_decomposed):
                _legs.extend([i] if c.ndim == 3 else [i, i])
            legs = tuple(_legs)
        if len(legs) != len(self.shape):
            raise ValueError(f"Tensor shape {self.shape} and legs {legs} are different")
        self.legs = tuple(legs)

    @property
    def dtype(self):
        return self.tensor_decomposed[0].dtype

    def decompose(self, **ignored) -> list[np.ndarray]:
        return self.tensor_decomposed

    def restore_from_decoposed(self) -> np.ndarray:
        """The dense operator tensor the cores represent, legs in site order (small operators only; the reference's name,
        typo included: dvr_operator_cls.py:547-554)."""
        t = self.tensor_decomposed[0]
        for core in self.tensor_decomposed[1:]:
            t = np.tensordot(t, core, axes=(-1, 0))
        return t[0, ..., 0]


def construct_kinetic_operator(dvr_prims, coefs=None, forms: str = "mpo") -> dict:
    """The kinetic energy operator as the ``{key: TensorOperator}`` dictionary ``TensorHamiltonian(kinetic=[[...]])`` takes
    (reference ``construct_kinetic_operator``, dvr_operator_cls.py:1135-1196): ``forms="mpo"`` = one key over all modes holding
    the bond-dimension-2 MPO of ``construct_kinetic_mpo``; ``forms="sop"`` = one single-site key per mode."""
    n = len(dvr_prims)
    coefs = [1.0] * n if coefs is None else list(coefs)
    if forms.lower() == "mpo":
        return {tuple((i, i) for i in range(n)): TensorOperator(mpo=construct_kinetic_mpo(dvr_prims, coefs))}
    if forms.lower() == "sop":
        out = {}
        for i, (prim, c) in enumerate(zip(dvr_prims, coefs, strict=True)):
            T = (-0.5 * c * prim.get_2nd_derivative_matrix_dvr()).astype(np.complex128)
            out[((i, i),)] = TensorOperator(mpo=[T.reshape(1, *T.shape, 1)], legs=(i, i))
        return out
    raise ValueError("forms must be 'sop' or 'mpo'")


def construct_kinetic_mpo(dvr_prims, coefs=None) -> list[np.ndarray]:
    """sum_i -1/2 c_i d^2/dQ_i^2 as a bond-dimension-2 MPO of full cores."""
    n = len(dvr_prims)
    coefs = [1.0] * n if coefs is None else list(coefs)
    mpo = []
    for i, (prim, c) in enumerate(zip(dvr_prims, coefs, strict=True)):
        g = prim.ngrid
        T = -0.5 * prim.get_2nd_derivative_matrix_dvr() * c
        eye = np.eye(g)
        if n == 1:
            W = np.zeros((1, g, g, 1), dtype=np.complex128)
            W[0, :, :, 0] = T
        elif i == 0:
            W = np.zeros((1, g, g, 2), dtype=np.complex128)
            W[0, :, :, 0] = T
            W[0, :, :, 1] = eye
        elif i == n - 1:
            W = np.zeros((2, g, g, 1), dtype=np.complex128)
            W[0, :, :, 0] = eye
            W[1, :, :, 0] = T
        else:
            W = np.zeros((2, g, g, 2), dtype=np.complex128)
            W[0, :, :, 0] = eye
            W[1, :, :, 0] = T
            W[1, :, :, 1] = eye
        mpo.append(W)
    return mpo
