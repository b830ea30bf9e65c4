"""Where does the HOST time of a launch-bound step go?  cProfile of K steps of a small-D bench workload (config 2 by default),
top functions by own time, next to the wall time per step and the same steps with the host work subtracted (an upper bound
of what removing Python overhead could give).  usage: python scripts/debug/host_profile.py [c2|c1] [steps]"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from pytdscf_b200 import workloads  # noqa: E402
from pytdscf_b200._const_cls import RunConfig  # noqa: E402
from pytdscf_b200._engine import Engine  # noqa: E402
from pytdscf_b200._mps_cuda import DeviceMPO, MPSCoefCuda  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = workloads.by_name(name)
eng = Engine(0)
model = wl.model()
H = DeviceMPO(eng, model.hamiltonian, merge_terms=True)
cfg = RunConfig(jobname="prof", space=wl.space, integrator=wl.integrator, conserve_norm=wl.conserve_norm)
mps = MPSCoefCuda.alloc_random(eng, model)
for _ in range(3):
    mps.propagate(wl.dt_au, H, cfg)
torch.cuda.synchronize()
l0 = eng.stats()["launches"]
t0 = time.perf_counter()
for _ in range(steps):
    mps.propagate(wl.dt_au, H, cfg)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / steps
launches = (eng.stats()["launches"] - l0) / steps
print(f"{wl.name}: {wall * 1e3:.1f} ms/step  {2 / wall:.2f} sweeps/s  {launches:.0f} launches/step  {wall * 1e6 / launches:.1f} us/launch")

# the same steps with the stream never waited for except where the algorithm reads a value: how long does the host alone need?
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    mps.propagate(wl.dt_au, H, cfg)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22)
print(s.getvalue()[:5000])
