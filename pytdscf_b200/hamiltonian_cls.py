"""``TensorHamiltonian``: the MPO Hamiltonian container (reference: ``pytdscf/hamiltonian_cls.py:618-752``).

``potential`` / ``kinetic`` are ``[[{key: TensorOperator}]]`` (or a bare dict for one state); keys are tuples
with ``int`` = diagonal leg, ``(k, k)`` = bra/ket legs, ``()`` = scalar -> ``coupleJ``.  ``backend`` may be
"numpy", "jax" (accepted for API compatibility; data stays NumPy until upload) or "cuda"."""
from __future__ import annotations

import itertools

import numpy as np

from ._mpo_cls import MatrixProductOperators, check_backend
from .dvr_operator_cls import TensorOperator


class TensorHamiltonian:
    def __init__(self, ndof: int, potential, name: str = "hamiltonian", kinetic=None, decompose_type: str = "QRD",
                 rate: float | None = None, bond_dimension=None, backend: str = "cuda"):
        backend = check_backend(backend)
        self.deferred = potential is None
        if self.deferred:
            # The reference's MPI scripts build the operators on rank 0 only and pass ``potential=None`` elsewhere
            # (tests/test_mpi_exiciton_propagate.py:175-183); its ``distribute_mpo_cores`` (hamiltonian_cls.py:805-850) then
            # scatters the cores.  Here such a placeholder adopts rank 0's cores when the site-parallel run starts
            # (``Simulator._distributed_wavefunction`` -> ``adopt``).
            self.name, self.nstate, self.ndof, self.backend = name, 1, ndof, backend
            self.coupleJ = [[0.0]]
            self.mpo = [[MatrixProductOperators(nsite=ndof, operators={}, backend=backend)]]
            return
        if isinstance(potential, dict):
            potential = [[potential]]
        if kinetic is not None and isinstance(kinetic, dict):
            kinetic = [[kinetic]]
        nstate = len(potential)
        self.name = name
        self.nstate = nstate
        self.ndof = ndof
        self.backend = backend
        self.coupleJ = [[0.0 for _ in range(nstate)] for _ in range(nstate)]
        self.mpo: list[list[MatrixProductOperators | None]] = [[None] * nstate for _ in range(nstate)]
        for i, j in itertools.product(range(nstate), repeat=2):
            operators: dict = {}
            if potential[i][j] is not None:
                for key, tensor in potential[i][j].items():
                    if key == ():
                        if not isinstance(tensor, (float, complex, int)):
                            raise ValueError(f"scalar term must be scalar but {tensor} is {type(tensor)}")
                        self.coupleJ[i][j] = tensor
                        continue
                    flat: tuple = ()
                    for k in key:
                        flat += k if type(k) is tuple else (k,)
                    if flat != tensor.legs:
                        raise ValueError(f"Given potential key {key} is not consistent with tensor legs {tensor.legs}")
                    if tensor.dtype not in (np.complex128, np.float64):
                        raise ValueError(f"core dtype must be complex128 or float64 but {tensor.dtype} is given")
                    operators[key] = list(tensor.decompose())
            if kinetic is not None and kinetic[i][j] is not None:
                for key, d2 in kinetic[i][j].items():
                    if key in operators:
                        raise ValueError(f"key {key} is already set in potential. Concatenate KEO and PEO or set KEO as SOP")
                    operators[key] = list(d2.decompose())
            self.mpo[i][j] = MatrixProductOperators(nsite=ndof, operators=operators, backend=backend)

    def export_cores(self) -> dict:
        """What a rank without operators needs to run: the scalar term and the cores of every MPO key (host arrays)."""
        if self.nstate != 1:
            raise NotImplementedError("only one state is supported")
        return {"coupleJ": complex(self.coupleJ[0][0]), "operators": {key: [np.asarray(c) for c in cores]
                                                                       for key, cores in self.mpo[0][0].operators.items()}}

    def adopt(self, payload: dict):
        """Fill a ``potential=None`` placeholder with the cores rank 0 exported."""
        self.coupleJ = [[payload["coupleJ"]]]
        self.mpo = [[MatrixProductOperators(nsite=self.ndof, operators=dict(payload["operators"]), backend=self.backend)]]
        self.deferred = False

    def apply_backend(self, backend: str):
        self.backend = check_backend(backend)
        for row in self.mpo:
            for m in row:
                if m is not None:
                    m.apply_backend(backend)

    def project_subspace(self, subspace_inds: dict[int, tuple[int, ...]]):
        """Keep only the listed physical indices of the given sites (Liouville sub-space projection,
        reference ``hamiltonian_cls.py:852-879``)."""
        assert len(self.mpo) == 1, "Only one state is supported"
        mpo = self.mpo[0][0]
        for isite, P in subspace_inds.items():
            for core in mpo.calc_point[isite]:
                if isinstance(core.data, int):
                    continue
                if core.data.ndim == 3:
                    core.data = core.data[:, P, :]
                else:
                    ket, bra = np.ix_(P, P)
                    core.data = core.data[:, ket, bra, :]
                core.shape, core.size = core.data.shape, core.data.size


__all__ = ["TensorHamiltonian", "TensorOperator"]
