"""Minimal opt_einsum (3.4 API subset) on top of np.einsum with an unlimited-memory greedy path.

Test infrastructure (oracle/refshim/README.md).  opt_einsum's default has no memory limit, so the
pairwise (BLAS-backed) contraction order must be forced the same way here.
"""
import numpy as np

_NOLIMIT = ("greedy", 2**62)


def contract(subscripts, *operands, **kwargs):
    path = np.einsum_path(subscripts, *operands, optimize=_NOLIMIT)[0]
    return np.einsum(subscripts, *operands, optimize=path)


class _Expression:
    def __init__(self, subscripts, shapes_or_consts, constants):
        self.subscripts = subscripts
        self.slots = list(shapes_or_consts)
        self.const_idx = sorted(constants)
        self.var_idx = [i for i in range(len(self.slots)) if i not in self.const_idx]
        dummies = [
            np.empty(s, dtype=np.complex128) if i in self.var_idx else np.asarray(s)
            for i, s in enumerate(self.slots)
        ]
        self.path = np.einsum_path(subscripts, *dummies, optimize=_NOLIMIT)[0]

    def __call__(self, *arrays, **kwargs):
        ops = list(self.slots)
        for i, a in zip(self.var_idx, arrays, strict=True):
            ops[i] = a
        return np.einsum(self.subscripts, *ops, optimize=self.path)


def contract_expression(subscripts, *shapes, constants=None, **kwargs):
    return _Expression(subscripts, shapes, constants or [])
