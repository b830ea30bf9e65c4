"""``Model`` / ``BasInfo``: wavefunction basis + operators (reference: ``pytdscf/model_cls.py:29-460``).

Scope of ``backend="cuda"``: MPS standard method (no SPF layer), MPO Hamiltonians, one "state" (electronic
levels are an exciton site), Hilbert or Liouville space."""
from __future__ import annotations

import copy

import numpy as np

from .dvr_operator_cls import TensorOperator
from .hamiltonian_cls import TensorHamiltonian


class BasInfo:
    def __init__(self, prim_info, spf_info=None, ndof_per_sites=None):
        if spf_info is not None or ndof_per_sites:
            raise NotImplementedError("the SPF (MCTDH) layer is deprecated in the reference and out of scope here")
        self.prim_info = copy.deepcopy(prim_info)
        self.is_standard_method = True
        self.is_DVR = any(hasattr(b, "get_grids") for b in prim_info[0])

    def get_nstate(self) -> int:
        return len(self.prim_info)

    def get_ndof(self) -> int:
        return len(self.prim_info[0])

    def get_primbas(self, istate: int, idof: int):
        return self.prim_info[istate][idof]

    def get_nprim(self, istate: int, idof: int) -> int:
        return self.prim_info[istate][idof].nprim

    get_ngrid = get_nprim
    get_nspf = get_nprim

    def get_nspf_list(self, istate: int) -> list[int]:
        return [self.get_nspf(istate, i) for i in range(self.get_ndof())]


class Model:
    init_weight_VIBSTATE: list | None = None
    init_weight_ESTATE: list | None = None
    init_HartreeProduct: list | None = None  # [state][dof][basis] weights or 3-D cores
    m_aux_max: int | None = None

    def __init__(self, basinfo, operators, *, bond_dim: int | None = None, build_td_hamiltonian=None,
                 space: str = "hilbert", subspace_inds: dict | None = None, one_gate_to_apply=None, kraus_op=None):
        if isinstance(basinfo, BasInfo):
            self.basinfo = basinfo
        elif isinstance(basinfo, list):
            self.basinfo = BasInfo(prim_info=basinfo if isinstance(basinfo[0], list) else [basinfo])
        else:
            raise TypeError("basinfo must be BasInfo instance or list.")
        if build_td_hamiltonian is not None:
            raise NotImplementedError("time-dependent Hamiltonians are not part of the MPO hot path")
        if kraus_op is not None:
            if not isinstance(kraus_op, dict) or any(len(k) not in (1, 2) for k in kraus_op):
                raise ValueError("kraus_op must be {(site,): B[k, d, d]} or {(site, site + 1): B[k, d, d]}")
        if one_gate_to_apply is not None and not isinstance(one_gate_to_apply, TensorHamiltonian):
            raise TypeError("one_gate_to_apply must be a TensorHamiltonian of one-site cores")
        if isinstance(operators, (TensorHamiltonian, list)):
            operators = {"hamiltonian": operators}
        ops = self.operators_to_tensor_hamiltonian(dict(operators))
        self.hamiltonian: TensorHamiltonian = ops.pop("hamiltonian")
        self.observables: dict[str, TensorHamiltonian] = ops
        if self.hamiltonian.nstate != self.basinfo.get_nstate():
            raise ValueError("The number of states in Hamiltonian and BasInfo are different.")
        if self.hamiltonian.nstate != 1:
            raise NotImplementedError("backend='cuda' supports nstate == 1 (electronic states as an exciton site)")
        self.nstate = 1
        self.m_aux_max = bond_dim
        self.use_mpo = True
        if space.lower() not in ("hilbert", "liouville"):
            raise ValueError(f"space must be 'hilbert' or 'liouville' but got {space}")
        self.space = space.lower()
        self.subspace_inds = None
        if self.space == "liouville" and subspace_inds is not None:
            self.subspace_inds = subspace_inds
            for op in self.observables.values():
                op.project_subspace(subspace_inds)
            self.hamiltonian.project_subspace(subspace_inds)
        self.one_gate_to_apply = one_gate_to_apply
        if one_gate_to_apply is not None and self.subspace_inds is not None:
            one_gate_to_apply.project_subspace(self.subspace_inds)
        self.kraus_op = None if kraus_op is None else {tuple(int(i) for i in k): np.asarray(v, dtype=np.complex128)
                                                       for k, v in kraus_op.items()}

    # -- basis passthroughs ------------------------------------------------------------------------
    def get_nstate(self) -> int:
        return self.basinfo.get_nstate()

    def get_ndof(self) -> int:
        return self.basinfo.get_ndof()

    def get_primbas(self, istate: int, idof: int):
        return self.basinfo.get_primbas(istate, idof)

    def get_nspf(self, istate: int, idof: int) -> int:
        return self.basinfo.get_nspf(istate, idof)

    def get_nprim(self, istate: int, idof: int) -> int:
        return self.basinfo.get_nprim(istate, idof)

    def get_nspf_list(self, istate: int) -> list[int]:
        return self.basinfo.get_nspf_list(istate)

    # -- operators ---------------------------------------------------------------------------------
    def guess_legkeys_from_mpo(self, mpo) -> tuple:
        if not isinstance(mpo, list):
            raise TypeError("mpo must be a list of arrays.")
        if len(mpo) != self.get_ndof():
            raise ValueError(f"mpo length must be equal to ndof of basis. But, got {len(mpo)} and {self.get_ndof()}.")
        key = []
        for k, core in enumerate(mpo):
            if core.ndim == 3:
                key.append((k,))
            elif core.ndim == 4:
                key.append((k, k))
            else:
                raise ValueError(f"Invalid core shape {core.shape} in mpo, {k}-th site.")
        return tuple(key)

    def operators_to_tensor_hamiltonian(self, operators: dict) -> dict:
        out: dict = {}
        if "potential" in operators:
            pot = operators.pop("potential")
            if isinstance(pot, TensorHamiltonian):
                raise ValueError("The 'potential' key must be list[np.ndarray] MPO.")
            if "hamiltonian" in operators:
                raise ValueError("Cannot specify 'hamiltonian' when 'potential' is given.")
            pot_key = self.guess_legkeys_from_mpo(pot)
            kin = operators.pop("kinetic", None)
            kinetic = None
            if kin is not None:
                if isinstance(kin, TensorHamiltonian):
                    raise ValueError("The 'kinetic' key must be list[np.ndarray] MPO.")
                kinetic = {self.guess_legkeys_from_mpo(kin): TensorOperator(mpo=kin)}
            out["hamiltonian"] = TensorHamiltonian(ndof=self.get_ndof(), potential={pot_key: TensorOperator(mpo=pot)},
                                                   kinetic=kinetic, backend="cuda")
        for name, op in operators.items():
            if isinstance(op, TensorHamiltonian):
                out[name] = op
            elif isinstance(op, list):
                if len(op) != self.get_ndof():
                    raise ValueError(f"Operator {name} length must be equal to ndof of basis. But, got {len(op)} and {self.get_ndof()}.")
                key = self.guess_legkeys_from_mpo(op)
                out[name] = TensorHamiltonian(ndof=self.get_ndof(), potential={key: TensorOperator(mpo=op)}, backend="cuda")
            else:
                raise TypeError(f"Operator {name} must be TensorHamiltonian or list of arrays.")
        return out

    def apply_backend(self, backend: str) -> None:
        self.hamiltonian.apply_backend(backend)
        for ob in self.observables.values():
            ob.apply_backend(backend)
        if self.one_gate_to_apply is not None:
            self.one_gate_to_apply.apply_backend(backend)

    # -- initial condition (reference: MPSCoef._get_initial_condition, _mps_cls.py:133-215) -----------
    def initial_core_weights(self) -> tuple[list, float, int]:
        """Per-site initial weights (1-D vectors or 3-D cores), the state weight scale and the bond cap."""
        ndof = self.get_ndof()
        m = 10**9 if self.m_aux_max is None else int(self.m_aux_max)
        if self.init_weight_ESTATE is not None:
            w = np.array(self.init_weight_ESTATE, dtype=float)
            if len(w) != 1 or w.min() < 0.0:
                raise ValueError("The length of weight_estate must be equal to nstate (1) and positive.")
        if self.init_HartreeProduct is not None:
            return list(self.init_HartreeProduct[0]), 1.0, m
        if self.init_weight_VIBSTATE is None:
            vib = [[1.0] + [0.0] * (self.get_nspf(0, i) - 1) for i in range(ndof)]
        else:
            if len(self.init_weight_VIBSTATE) != 1:
                raise ValueError("The length of weight_vib must be equal to nstate.")
            vib = self.init_weight_VIBSTATE[0]
            if len(vib) != ndof:
                raise ValueError(f"The length of weight_vib[0] must be equal to ndof But len(weight_vib[0]) = {len(vib)} != {ndof}")
        # HO-DVR sites: FBR weights are rotated into the DVR representation (_mps_cls.py:2669-2682).
        # The reference normalises the FBR vector first (init_random) and only then applies the unitary.
        cores = []
        for i in range(ndof):
            prim = self.get_primbas(0, i)
            v = np.array(vib[i], dtype=np.complex128)
            if hasattr(prim, "get_unitary"):
                if self.space == "hilbert":
                    v = v / np.linalg.norm(v)
                else:
                    import math

                    q = math.isqrt(len(v))
                    v = v / np.trace(v.reshape(q, q))
                v = v @ prim.get_unitary()
                cores.append(v.reshape(1, -1, 1))  # a 3-D core is copied verbatim (no re-normalisation)
            else:
                cores.append(v)
        return cores, 1.0, m
