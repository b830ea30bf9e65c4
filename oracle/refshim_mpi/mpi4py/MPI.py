import os
import pickle
import sys
import time

LOR = "LOR"
SUM = "SUM"
MAX = "MAX"


class _Comm:
    def __init__(self):
        self.rank = int(os.environ["FAKE_MPI_RANK"])
        self.size = int(os.environ["FAKE_MPI_SIZE"])
        self.dir = os.environ["FAKE_MPI_DIR"]
        self._seq = {}
        self._bar = 0
        self._coll = 0

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def _path(self, src, dst, tag, seq):
        return os.path.join(self.dir, f"m_{src}_{dst}_{tag}_{seq}.pkl")

    def send(self, obj, dest, tag=0):
        key = ("s", dest, tag)
        seq = self._seq.get(key, 0)
        self._seq[key] = seq + 1
        p = self._path(self.rank, dest, tag, seq)
        with open(p + ".tmp", "wb") as f:
            pickle.dump(obj, f)
        os.rename(p + ".tmp", p)

    def recv(self, source, tag=0):
        key = ("r", source, tag)
        seq = self._seq.get(key, 0)
        self._seq[key] = seq + 1
        p = self._path(source, self.rank, tag, seq)
        t0 = time.time()
        while not os.path.exists(p):
            if os.path.exists(os.path.join(self.dir, "ABORT")) or time.time() - t0 > 600:
                sys.exit(1)
            time.sleep(0.0005)
        with open(p, "rb") as f:
            obj = pickle.load(f)
        os.remove(p)
        return obj

    def barrier(self):
        n = self._bar
        self._bar += 1
        open(os.path.join(self.dir, f"b_{n}_{self.rank}"), "w").close()
        t0 = time.time()
        for r in range(self.size):
            while not os.path.exists(os.path.join(self.dir, f"b_{n}_{r}")):
                if os.path.exists(os.path.join(self.dir, "ABORT")) or time.time() - t0 > 600:
                    sys.exit(1)
                time.sleep(0.0005)

    Barrier = barrier

    def _ctag(self):
        self._coll += 1
        return 100000 + self._coll

    def bcast(self, obj, root=0):
        tag = self._ctag()
        if self.rank == root:
            for r in range(self.size):
                if r != root:
                    self.send(obj, r, tag)
            return obj
        return self.recv(root, tag)

    def scatter(self, objs, root=0):
        tag = self._ctag()
        if self.rank == root:
            for r in range(self.size):
                if r != root:
                    self.send(objs[r], r, tag)
            return objs[root]
        return self.recv(root, tag)

    def gather(self, obj, root=0):
        tag = self._ctag()
        if self.rank == root:
            out = [None] * self.size
            out[root] = obj
            for r in range(self.size):
                if r != root:
                    out[r] = self.recv(r, tag)
            return out
        self.send(obj, root, tag)
        return None

    def allgather(self, obj):
        return self.bcast(self.gather(obj, 0), 0)

    def allreduce(self, obj, op=SUM):
        vals = self.allgather(obj)
        if op == LOR:
            return any(vals)
        if op == MAX:
            return max(vals)
        return sum(vals)

    def Abort(self, code=1):
        open(os.path.join(self.dir, "ABORT"), "w").close()
        sys.exit(code)


COMM_WORLD = _Comm()
