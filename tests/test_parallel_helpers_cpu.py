"""CPU unit tests of the host-side helpers of the site-parallel path: ``balanced_split`` (choice of
``parallel_split_indices``, pytdscf/simulator_cls.py:178 takes them from the user) and the tagged message protocol of
``Comm`` (the replacement of the reference's pickled MPI messages, pytdscf/_mps_parallel.py:541-807), through a loop-back
stand-in for ``torch.distributed`` (the real gloo / NCCL paths are covered by tests/test_site_parallel_cpu.py and the GPU tests)."""
import itertools
import types

import numpy as np
import pytest
import torch

from pytdscf_b200._mps_cuda import Block, SiteCoef, bond_dims
from pytdscf_b200._mps_parallel import Comm, balanced_split


def _costs(dims, D):
    out = []
    for i, d in enumerate(dims):
        dl, dr = bond_dims(dims, i, D)
        out.append(float(d) * (dl * dl * dr + dl * dr * dr))
    return out


def _worst(costs, starts):
    ends = list(starts[1:]) + [len(costs)]
    return max(sum(costs[a:b]) for a, b in zip(starts, ends, strict=True))


@pytest.mark.parametrize("dims,D,nranks", [([4] * 12, 16, 3), ([2, 3, 4, 5, 4, 3, 2, 2, 6, 2], 8, 2), ([8] * 16, 64, 4), ([3] * 9, 5, 4)])
def test_balanced_split_is_the_min_max_partition(dims, D, nranks):
    starts = balanced_split(dims, D, nranks)
    n = len(dims)
    assert starts[0] == 0 and len(starts) == nranks and starts == sorted(starts)
    ends = starts[1:] + [n]
    assert all(b - a >= 2 for a, b in zip(starts, ends, strict=True))          # every segment holds two sites (joint boundary update)
    costs = _costs(dims, D)
    best = min(_worst(costs, [0, *cut]) for cut in itertools.combinations(range(2, n - 1), nranks - 1)
               if all(b - a >= 2 for a, b in zip([0, *cut], [*cut, n], strict=True)))
    assert _worst(costs, starts) == pytest.approx(best, rel=1e-12)


def test_balanced_split_gives_the_chain_ends_more_sites():
    # small bonds near the ends are nearly free: the end segments of config 5 (128 sites, D = 512) are the longest
    starts = balanced_split([8] * 128, 512, 8)
    lens = [b - a for a, b in zip(starts, starts[1:] + [128], strict=True)]
    assert lens[0] > lens[3] and lens[-1] > lens[4] and sum(lens) == 128


def test_balanced_split_rejects_too_many_ranks():
    with pytest.raises(ValueError):
        balanced_split([4] * 6, 8, 4)
    assert balanced_split([4] * 6, 8, 1) == [0]


# ---------------------------------------------------------------------------------------------------------
class _LoopDist:
    """Just enough of torch.distributed for Comm: messages are queued in order and delivered to the same process."""

    def __init__(self):
        self.objects: list = []
        self.tensors: list = []
        self.headers = 0
        self.batches = 0
        self.isend, self.irecv = "isend", "irecv"

    def get_backend(self):
        return "loop"

    def send_object_list(self, objs, dst):
        self.headers += 1
        self.objects.append(objs[0])

    def recv_object_list(self, box, src):
        box[0] = self.objects.pop(0)

    def P2POp(self, op, tensor, peer):
        return (op, tensor, peer)

    def batch_isend_irecv(self, ops):
        self.batches += 1
        for op, tensor, _peer in ops:
            if op == "isend":
                self.tensors.append(tensor.clone())
            else:
                tensor.copy_(self.tensors.pop(0))
        return [types.SimpleNamespace(wait=lambda: None)]


def _comm():
    dist = _LoopDist()
    info = types.SimpleNamespace(dist=dist, rank=0, world=2)
    return Comm(info, torch.device("cpu")), dist


def _crand(rng, *shape):
    return torch.from_numpy(rng.standard_normal(shape) + 1j * rng.standard_normal(shape))


def test_comm_round_trip_of_nested_containers():
    comm, dist = _comm()
    rng = np.random.default_rng(3)
    msg = {"blocks": {("a", 1): Block(_crand(rng, 3, 2, 3), False, 3), ("ovlp",): Block(_crand(rng, 3, 3), True, 3)},
           "site": SiteCoef(_crand(rng, 3, 4, 5), "Psi", 7), "pair": (_crand(rng, 2, 2), [1.5, "B", None])}
    comm.send(msg, 1)
    got = comm.recv(1)
    assert set(got) == set(msg) and got["site"].gauge == "Psi" and got["site"].isite == 7
    assert torch.equal(got["site"].data, msg["site"].data)
    for k, b in msg["blocks"].items():
        assert torch.equal(got["blocks"][k].data, b.data) and got["blocks"][k].is_identity == b.is_identity
    assert torch.equal(got["pair"][0], msg["pair"][0]) and got["pair"][1] == [1.5, "B", None] and isinstance(got["pair"], tuple)


def test_comm_tag_sends_the_layout_once_and_one_batch_per_message():
    comm, dist = _comm()
    rng = np.random.default_rng(4)
    for step in range(3):
        msg = {"L": _crand(rng, 4, 2, 4), "x": _crand(rng, 4, 4)}
        comm.send(msg, 1, tag="1a")
        got = comm.recv(1, tag="1a")
        assert torch.equal(got["L"], msg["L"]) and torch.equal(got["x"], msg["x"])
    assert dist.headers == 1                   # the container description travelled with the first message only
    assert dist.batches == 6                   # one batched p2p launch per send and per receive, whatever the tensor count
    comm.send({"L": _crand(rng, 4, 2, 4), "x": _crand(rng, 4, 4)}, 1)      # untagged: self-describing every time
    comm.recv(1)
    assert dist.headers == 2


def test_comm_refuses_a_changed_layout_under_the_same_tag():
    comm, _ = _comm()
    rng = np.random.default_rng(5)
    comm.send({"x": _crand(rng, 4, 4)}, 1, tag="2c")
    with pytest.raises(RuntimeError, match="layout"):
        comm.send({"x": _crand(rng, 5, 4)}, 1, tag="2c")      # the receiver would misread the bytes
    with pytest.raises(RuntimeError, match="layout"):
        comm.send({"y": _crand(rng, 4, 4)}, 1, tag="2c")
