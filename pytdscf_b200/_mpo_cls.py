"""Per-site MPO term bookkeeping (reference: ``MatrixProductOperators`` / ``OperatorCore``,
``pytdscf/_mpo_cls.py:44-234``).  Pure host-side metadata; core data moves to HBM once, in
``MPSCoefCuda.upload_hamiltonian``."""
from __future__ import annotations

import numpy as np

BACKENDS = ("numpy", "jax", "cuda")


def check_backend(backend: str) -> str:
    b = backend.lower()
    if b not in BACKENDS:
        raise ValueError(f"backend must be jax, numpy or cuda, but {backend} is given")
    return b


class OperatorCore:
    """One core W of one MPO term acting at ``psite``; ``data`` is ``1`` (int) for an identity gap core."""

    def __init__(self, parent_key: list[int], original_key: tuple, psite: int, data, backend: str):
        self.key = original_key
        self.psite = psite
        self.data = data
        self.is_right_side = max(parent_key) == psite
        self.is_left_side = min(parent_key) == psite
        self.is_unitmat_op = isinstance(data, int)
        self.only_diag = isinstance(data, int) or data.ndim == 3
        self.backend = backend
        if not isinstance(data, int):
            self.shape = data.shape
            self.size = data.size

    def apply_backend(self, backend: str):
        self.backend = check_backend(backend)

    def __repr__(self) -> str:
        return f"OperatorCore(key={self.key}, site={self.psite}, diag={self.only_diag})"


class MatrixProductOperators:
    """``operators``: ``{key: [cores]}``; ``calc_point[p]`` lists the cores acting at site p."""

    def __init__(self, nsite: int, operators: dict, backend: str):
        self.nsite = nsite
        self.operators = operators
        self.backend = check_backend(backend)
        self.calc_point: list[list[OperatorCore]] = [[] for _ in range(nsite)]
        for key, mpo in operators.items():
            sites: list[int] = []
            for ind, core in zip(key, mpo, strict=True):
                if isinstance(ind, tuple):
                    if len(ind) != core.ndim - 2:
                        raise ValueError(f"MPO core index {ind} is not consistent with MPO core shape {core.shape}")
                    if len(set(ind)) != 1:
                        raise ValueError(f"MPO core index must be single DOFs, but assinged {ind}")
                    sites.append(ind[0])
                elif isinstance(ind, (int, np.integer)):
                    if core.ndim != 3:
                        raise ValueError(f"MPO core index {ind} in {key} is not consistent with MPO core shape {core.shape}")
                    sites.append(int(ind))
                else:
                    raise ValueError(f"MPO core index type is wrong. Must be int or Tuple[int,int], but {ind} is given")
            for s, core in zip(sites, mpo, strict=True):
                self.calc_point[s].append(OperatorCore(sites, key, s, core, self.backend))
            for s in range(min(sites) + 1, max(sites)):
                if s not in sites:
                    self.calc_point[s].append(OperatorCore(sites, key, s, 1, self.backend))

    def apply_backend(self, backend: str):
        self.backend = check_backend(backend)
        for cores in self.calc_point:
            for c in cores:
                c.apply_backend(backend)
