#!/usr/bin/env python
"""Benchmark of the TDVP hot path (BASELINE.json metric: TDVP sweeps/s and H_eff contraction TFLOP/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c3|c2|c5] [--impl cuda|reference]

A "step" is one TDVP time step = one forward + one backward half sweep (2 sweeps) of one-site
projector-splitting TDVP over the whole chain (``MPSCoef.propagate``, pytdscf/_mps_cls.py:452-503), properties
off.  Default workload: BASELINE config 4 (radical-pair Liouville-space MPDO, D = 1024, Arnoldi), the
configuration the north-star target (>= 20x the CPU path at D = 1024) is quoted on.

* ``value``   sweeps/s with the MPS, MPO and environments resident in HBM (CUDA events, max over ranks).
* ``e2e``     same metric through the public host-buffer path: every step copies the site tensors from pinned host
              memory (H2D), rebuilds the environments, propagates, and copies the tensors + the norm back (D2H).
* ``roofline``  the dominant kernel (zgemm_dmma_kernel): sum of 8*M*N*K over its launches in the timed region /
              sum of their CUDA-event durations on the launching stream, against the measured FP64 DMMA peak
              (launch-bound workloads, D <= 64: the event pairs move to a second pass of the same steps, see
              ``roofline.measured_in``).
* ``cpu_baseline``  the oracle's NumPy/BLAS restatement of the reference path, timed on this box's host cores on a
              bounded sample (one full-bond-dimension site update), extrapolated to sweeps/s by algorithmic flops.
* N > 1       config 4 is "replicas only" in the reference's semantics (Liouville space is excluded from
              its parallel path; trajectories / initial states are independent): N independent replicas, no
              data-path collective, weak scaling.  ``--workload c5`` (config 5, 128 sites, D = 512) runs the reference's
              site-segment-parallel TDVP of ONE chain over NCCL point-to-point instead (strong scaling;
              ``--parallel replicas|sites`` overrides the choice).
``--impl reference`` times the CPU path alone (rank 0), each step being the bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP64_DMMA_PEAK_TFLOPS = 37.0  # measured on this pool's B200: profiles/r1_fp64_pipe_microbench.jsonl (64 FMA/clk/SM @1.965 GHz)
STATS_DIR = os.path.join(ROOT, "bench_stats")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--bond-dim", type=int, default=None, help="override the workload's bond dimension (not a bench line)")
    ap.add_argument("--parallel", default="auto", choices=["auto", "replicas", "sites"],
                    help="N > 1: 'sites' = site-segment-parallel TDVP of ONE chain (strong scaling; default for c5), "
                         "'replicas' = N independent chains (weak scaling; default otherwise)")
    ap.add_argument("--sites", type=int, default=None, help="override the chain length of c5 (not a bench line)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def make_workload(args):
    from pytdscf_b200 import workloads

    kw = {}
    if args.bond_dim is not None:
        kw["D"] = args.bond_dim
    if getattr(args, "sites", None) is not None and args.workload == "c5":
        kw["nsite"] = args.sites
    return workloads.by_name(args.workload, **kw)


# ----------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows: list[list[str]] = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
# CPU sample: one site update of the reference algorithm (oracle port) on this box's host cores
# ----------------------------------------------------------------------------------------------------
def _herm_block(rng, D, w):
    x = (rng.standard_normal((D, w, D)) + 1j * rng.standard_normal((D, w, D))) / np.sqrt(2 * D)
    return (x + x.conj().transpose(2, 1, 0)) / 2


def site_shapes(wl):
    from pytdscf_b200._mps_cuda import bond_dims

    return [(bond_dims(wl.dims, i, wl.bond_dim), wl.dims[i]) for i in range(len(wl.dims))]


def pick_sample_site(wl) -> int:
    """A full-bond-dimension site of the most common physical dimension (representative of the sweep's bulk)."""
    shapes = site_shapes(wl)
    full = [i for i, ((dl, dr), d) in enumerate(shapes) if dl == wl.bond_dim and dr == wl.bond_dim] or \
           [int(np.argmax([dl * dr for (dl, dr), _ in shapes]))]
    dcount: dict = {}
    for i in full:
        dcount[wl.dims[i]] = dcount.get(wl.dims[i], 0) + 1
    dmode = max(dcount, key=dcount.get)
    cand = [i for i in full if wl.dims[i] == dmode]
    return cand[len(cand) // 2]


def cpu_site_update_sample(wl, n_H: int, n_K: int, seed: int = 7):
    """Time n_H H_eff applications + QR shift + environment update + n_K K_eff applications + absorb at one
    full-D site with the oracle's kernels (NumPy einsum on a BLAS-backed greedy path + SciPy LAPACK, i.e. what
    the reference's NumPy backend executes).  Returns (seconds, algorithmic contraction flops)."""
    from oracle import tdvp_oracle as orc

    rng = np.random.default_rng(seed)
    p = pick_sample_site(wl)
    (Dl, Dr), d = site_shapes(wl)[p]
    H = orc.MPOHamiltonian(len(wl.dims), wl.operators, wl.coupleJ)
    psi = rng.standard_normal((Dl, d, Dr)) + 1j * rng.standard_normal((Dl, d, Dr))
    psi /= np.linalg.norm(psi)
    hterms, flops_H, flops_E, flops_K = {}, 0.0, 0.0, 0.0
    kterms = {}
    for core in H.calc_point[p]:
        wl_, wr_ = core.data.shape[0], core.data.shape[-1]
        L = _herm_block(rng, Dl, wl_) if not core.is_left else None
        R = _herm_block(rng, Dr, wr_) if not core.is_right else None
        hterms[core.key] = (L, core, R)
        full = core.data.ndim == 4
        if L is not None:
            flops_H += 8.0 * Dl * Dl * Dr * d * wl_
            flops_E += 8.0 * Dl * Dl * Dr * d * wl_
        flops_H += 8.0 * Dl * Dr * (d * d if full else d) * wl_ * wr_
        flops_E += 8.0 * Dl * Dr * (d * d if full else d) * wl_ * wr_ + 8.0 * Dl * Dr * Dr * d * wr_
        if R is not None:
            flops_H += 8.0 * Dl * Dr * Dr * d * wr_
            kterms[core.key] = (_herm_block(rng, Dr, wr_), R)
            flops_K += 16.0 * wr_ * Dr**3
    t0 = time.perf_counter()
    x = psi
    for _ in range(n_H):
        x = orc.heff_apply(hterms, 0.0, x)
        x /= np.linalg.norm(x)
    A, sigma = orc.shift_qr(x)
    for key, (L, core, R) in hterms.items():
        orc.env_update_term("A", A, A, L, core)
    s = sigma
    for _ in range(n_K):
        if kterms:
            s = orc.keff_apply(kterms, 0.0, s)
            s /= np.linalg.norm(s)
    np.tensordot(s, psi, axes=(1, 0))
    dt = time.perf_counter() - t0
    return dt, n_H * flops_H + flops_E + n_K * flops_K, {"site": p, "Dl": Dl, "d": d, "Dr": Dr, "n_H": n_H, "n_K": n_K}


def load_stats(wl_name: str) -> dict | None:
    path = os.path.join(STATS_DIR, wl_name + ".json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info

        return max([int(p.get("num_threads", 1)) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline(wl, stats: dict, repeats: int = 1) -> dict:
    n_H = max(1, round(stats["avg_matvecs_H"]))
    n_K = max(1, round(stats["avg_matvecs_K"]))
    best = None
    nrun = 0
    while nrun < repeats or (best[0] < 2.0 and nrun < 5):
        # samples shorter than 2 s (D <= 64 workloads) are repeated and the fastest kept: the first pass pays the
        # BLAS thread-pool start-up and einsum path search that a long reference run amortises
        sec, flops, info = cpu_site_update_sample(wl, n_H, n_K)
        nrun += 1
        if best is None or sec < best[0]:
            best = (sec, flops, info)
    sec, flops, info = best
    rate = flops / sec
    sweeps_per_s = rate / stats["flops_per_sweep"]
    return {"value": sweeps_per_s, "unit": "sweeps/s", "cores": blas_threads(), "kind": "port",
            "sample": (f"one site update at site {info['site']} (Dl={info['Dl']}, d={info['d']}, Dr={info['Dr']}): "
                       f"{n_H} H_eff applies + QR + env update + {n_K} K_eff applies + absorb on seeded random blocks, "
                       f"{sec:.2f} s, {rate / 1e9:.1f} GFLOP/s algorithmic; extrapolated with "
                       f"{stats['flops_per_sweep'] / 1e12:.3f} TFLOP/sweep"),
            "seconds": sec, "gflops": rate / 1e9}


# ----------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args)
    stats = load_stats(wl.name) or default_stats(wl)
    vals = []
    total = args.warmup + args.steps
    for i in range(total):
        b = cpu_baseline(wl, stats)
        if i >= args.warmup:
            vals.append(b)
    value = statistics.mean(v["value"] for v in vals)
    sec = statistics.mean(v["seconds"] for v in vals)
    cb = dict(vals[-1])
    cb["value"] = value
    out = {"impl": "reference", "metric": "tdvp_sweeps_per_sec", "value": value, "unit": "sweeps/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
           "config": {"workload": wl.name, "description": wl.description, "step": "bounded CPU sample (see cpu_baseline.sample)"},
           "cpu_baseline": cb,
           "e2e": {"value": value, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def default_stats(wl) -> dict:
    """Analytic stand-in when no GPU run has recorded the Krylov counts yet: 7 / 4 matvecs per H / K solve."""
    from oracle import tdvp_oracle as orc

    H = orc.MPOHamiltonian(len(wl.dims), wl.operators, wl.coupleJ)
    n_H, n_K = 7.0, 4.0
    flops = 0.0
    shapes = site_shapes(wl)
    n = len(wl.dims)
    for p, ((Dl, Dr), d) in enumerate(shapes):
        for core in H.calc_point[p]:
            wl_, wr_ = core.data.shape[0], core.data.shape[-1]
            full = core.data.ndim == 4
            fH = 8.0 * Dl * Dr * (d * d if full else d) * wl_ * wr_
            if not core.is_left:
                fH += 8.0 * Dl * Dl * Dr * d * wl_
            if not core.is_right:
                fH += 8.0 * Dl * Dr * Dr * d * wr_
            flops += n_H * fH * (2 if p in (0, n - 1) else 1) * 1.0 + fH  # + env update
            if not core.is_right:
                flops += n_K * 16.0 * wr_ * Dr**3
    return {"avg_matvecs_H": n_H, "avg_matvecs_K": n_K, "flops_per_sweep": flops, "source": "analytic default"}


# ----------------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl cuda needs a CUDA device (no CPU fallback)")
    from pytdscf_b200 import parallel

    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    info = parallel.init_from_env("nccl")
    rank, world, local_rank, dist = info.rank, info.world, info.local_rank, info.dist

    from pytdscf_b200._const_cls import RunConfig
    from pytdscf_b200._engine import Engine
    from pytdscf_b200._mps_cuda import DeviceMPO, MPSCoefCuda

    wl = make_workload(args)
    model = wl.model()
    eng = Engine(local_rank)
    H = DeviceMPO(eng, model.hamiltonian)
    cfg = RunConfig(jobname="bench", space=wl.space, integrator=wl.integrator, conserve_norm=wl.conserve_norm)
    site_parallel = world > 1 and (args.parallel == "sites" or (args.parallel == "auto" and args.workload == "c5"))
    if site_parallel:
        # one chain, contiguous site segments, one per GPU (reference: MPSCoefParallel, _mps_parallel.py:106-268)
        from pytdscf_b200._mps_parallel import Comm, MPSCoefParallelCuda

        n = len(wl.dims)
        split = [(r * n) // world for r in range(world)]
        mps = MPSCoefParallelCuda.distribute(eng, Comm(info, eng.torch_device), model, split)
    else:
        mps = MPSCoefCuda.alloc_random(eng, model)
    dt = wl.dt_au

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    us_per_launch = 1e9
    for iw in range(args.warmup):
        if iw == args.warmup - 1:
            torch.cuda.synchronize()
            l0, t0 = eng.stats()["launches"], time.perf_counter()
        mps.propagate(dt, H, cfg)
        if iw == args.warmup - 1:
            torch.cuda.synchronize()
            us_per_launch = (time.perf_counter() - t0) * 1e6 / max(1, eng.stats()["launches"] - l0)
    barrier()
    # Per-launch CUDA events (for the roofline) ride inside the timed region when the step is GPU-bound (>= 40 us of
    # step time per launch: two event records per launch are < 1 % there) and move to a separate pass of the same K
    # steps when the stream is launch-bound (D <= 64 workloads), where they would cost ~15 % of the reported time.
    profile_in_timed_region = us_per_launch >= 40.0
    if dist is not None:
        flag = torch.tensor([1.0 if profile_in_timed_region else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        profile_in_timed_region = bool(flag.item() > 0.5)

    # ---- timed region: K steps, state resident in HBM ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    eng.reset_stats()
    if profile_in_timed_region:
        eng.gemm_profile(True, reset=True)
    launches0 = eng.stats()["launches"]
    mps.record_trace = True
    mps.trace = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        mps.propagate(dt, H, cfg)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    if profile_in_timed_region:
        prof = eng.gemm_profile(False)
        breakdown = eng.profile_breakdown()
        ms_prof = ms
    st = eng.stats()
    launches = st["launches"] - launches0
    trace = np.array(mps.trace)
    mps.record_trace = False
    # ---- launch-bound workloads only: profiled pass (NOT the reported time), the same K steps again with a CUDA-event
    # pair around every kernel launch of the library
    if not profile_in_timed_region:
        eng.gemm_profile(True, reset=True)
        pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        pv0.record()
        for _ in range(args.steps):
            mps.propagate(dt, H, cfg)
        pv1.record()
        barrier()
        ms_prof = pv0.elapsed_time(pv1)
        prof = eng.gemm_profile(False)
        breakdown = eng.profile_breakdown()
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    jobs = 1 if site_parallel else world      # chains propagated by the whole job
    sweeps = 2 * args.steps * jobs
    value = sweeps / (ms * 1e-3)
    nH = trace[trace[:, 0] == 0][:, 2]
    nK = trace[trace[:, 0] == 1][:, 2]
    flops_all = st["flops"]
    if site_parallel:
        flops_all = parallel.sum_over_ranks(info, flops_all, device="cuda")
    flops_per_sweep = flops_all / (2 * args.steps)
    stats = {"avg_matvecs_H": float(nH.mean()), "avg_matvecs_K": float(nK.mean()) if len(nK) else 0.0,
             "flops_per_sweep": flops_per_sweep, "source": "bench.py GPU run"}
    if rank == 0 and not site_parallel and not os.path.exists(os.path.join(STATS_DIR, wl.name + ".json")):
        # the committed files (Krylov counts of the named workloads) keep the CPU legs of both arms on the same sample;
        # only workloads without one (e.g. --bond-dim overrides) record theirs here
        os.makedirs(STATS_DIR, exist_ok=True)
        with open(os.path.join(STATS_DIR, wl.name + ".json"), "w") as f:
            json.dump(stats, f, indent=1)

    # ---- e2e: host buffers in, host buffers out, every step ----
    e2e = None
    if not args.no_e2e and site_parallel:
        # segment state (site tensors, both saved gauges, boundary bond matrices) lives in pinned host memory between
        # steps; the environment blocks are caches derived from it and stay on the device
        def state_refs():
            refs = [(sc, "data") for grp in (mps.sites, mps.superblock_all_A, mps.superblock_all_B) for sc in grp]
            if mps.joint_sigvec is not None:
                refs += [(mps, "joint_sigvec"), (mps, "joint_sigvec_not_pinv")]
            return refs

        host = [torch.empty(getattr(o, a).shape, dtype=torch.complex128).pin_memory() for o, a in state_refs()]
        for hb, (o, a) in zip(host, state_refs(), strict=True):
            hb.copy_(getattr(o, a))
        nbytes = sum(hb.numel() * 16 for hb in host)
        e_steps = max(1, min(args.steps, 2))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e_steps):
            for hb, (o, a) in zip(host, state_refs(), strict=True):
                setattr(o, a, hb.to(eng.torch_device, non_blocking=True))
            mps.propagate(dt, H, cfg)
            refs = state_refs()
            if len(refs) != len(host) or any(getattr(o, a).shape != hb.shape for hb, (o, a) in zip(host, refs)):
                raise RuntimeError("segment state changed shape during a step")
            for hb, (o, a) in zip(host, refs, strict=True):
                hb.copy_(getattr(o, a), non_blocking=True)
        e1.record()
        barrier()
        ems = e0.elapsed_time(e1)
        t = torch.tensor([ems, float(nbytes)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
        ems, tot = float(t[0].item()), int(t[1].item())
        e2e = {"value": 2 * e_steps / (ems * 1e-3), "unit": "sweeps/s", "h2d_bytes_per_step": tot, "d2h_bytes_per_step": tot,
               "steps": e_steps, "path": "pinned host segment state -> H2D -> MPSCoefParallelCuda.propagate -> D2H, every rank"}
    elif not args.no_e2e:
        host = [torch.empty(s.data.shape, dtype=torch.complex128).pin_memory() for s in mps.sites]
        for hbuf, s in zip(host, mps.sites, strict=True):
            hbuf.copy_(s.data)
        gauges = [s.gauge for s in mps.sites]
        nbytes = sum(hb.numel() * 16 for hb in host)
        norm_host = torch.empty(1, dtype=torch.float64).pin_memory()
        e_steps = max(1, min(args.steps, 2))

        def e2e_step():
            cores = [hb.to(eng.torch_device, non_blocking=True) for hb in host]
            m2 = MPSCoefCuda(eng, cores, gauges)  # fresh object: environments are rebuilt from scratch
            m2.niter_krylov = dict(mps.niter_krylov)
            m2.propagate(dt, H, cfg)
            for hb, s in zip(host, m2.sites, strict=True):
                hb.copy_(s.data, non_blocking=True)
            norm_host.copy_(torch.linalg.vector_norm(m2.sites[0].data).reshape(1), non_blocking=True)
            mps.niter_krylov = m2.niter_krylov

        e2e_step()   # untimed: first use of the pinned buffers / torch's norm kernel initialises lazily
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e_steps):
            e2e_step()
        e1.record()
        barrier()
        ems = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ems], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": 2 * e_steps * world / (ems * 1e-3), "unit": "sweeps/s", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes + 8, "steps": e_steps,
               "path": "pinned host MPS -> H2D -> MPSCoefCuda.propagate (environments rebuilt) -> D2H of MPS + norm"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    gemm_tflops = prof["flops"] / (prof["ms"] * 1e-3) / 1e12 if prof["ms"] > 0 else 0.0
    ncu = None
    ncu_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(ncu_path):
        with open(ncu_path) as f:
            ncu = json.load(f).get(wl.name)
    out = {
        "metric": "tdvp_sweeps_per_sec", "value": value, "unit": "sweeps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if site_parallel else "weak",
        "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
        "config": {"workload": wl.name, "description": wl.description, "sites": len(wl.dims), "bond_dim": wl.bond_dim,
                   "integrator": wl.integrator, "dt_au": wl.dt_au, "thresh_sil": cfg.thresh_exp,
                   "step": "1 time step = 2 half sweeps, properties off",
                   "parallelism": "single GPU" if world == 1 else
                   (f"site-parallel TDVP, {world} contiguous segments, NCCL p2p of boundary blocks" if site_parallel
                    else f"{world} independent replicas (no collective)"),
                   "l2": "per-step working set (Krylov basis + contraction intermediates, GBs) >> 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "heff_tflops": flops_all / (ms * 1e-3) / 1e12,
        "heff_tflops_note": "algorithmic H_eff + K_eff + env-update flops of the reference's contraction (SURVEY 8(d) formula, identity "
                            "MPO channels included) / wall time of the timed region (all kernels); the roofline entry counts "
                            "the flops the GEMM launches actually execute",
        "krylov": {"avg_matvecs_H": stats["avg_matvecs_H"], "avg_matvecs_K": stats["avg_matvecs_K"],
                   "solves_per_step": len(trace) / args.steps, "tflop_per_sweep": flops_per_sweep / 1e12},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "zgemm_dmma_kernel", "bound": "tensor", "achieved": gemm_tflops, "peak": FP64_DMMA_PEAK_TFLOPS,
                     "unit": "TFLOP/s", "frac": gemm_tflops / FP64_DMMA_PEAK_TFLOPS,
                     "traffic": None if ncu is None else ncu.get("dram_bytes_per_launch"),
                     "launches": int(prof["launches"]), "share_of_step": prof["ms"] / ms_prof,
                     "measured_in": "the timed region" if profile_in_timed_region else
                     f"separate profiled pass of the same {args.steps} steps ({ms_prof / args.steps:.1f} ms/step with per-launch "
                     f"events; the step is launch-bound at {us_per_launch:.0f} us per launch)",
                     "peak_source": "measured FP64 DMMA peak of this pool's B200 (profiles/r1_fp64_pipe_microbench.jsonl; "
                                    "MEASURED_PEAKS.json has no FP64 entry; cuBLAS ZGEMM 8192^3 = 36.97 TFLOP/s)"},
    }
    # per-label CUDA-event breakdown of the timed region (share of the step per kernel family)
    top = sorted(breakdown.items(), key=lambda kv: -kv[1]["ms"])
    out["breakdown_top"] = {k: {"share": round(v["ms"] / ms_prof, 4), "launches": v["launches"],
                                "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] > 0 and v["flops"] > 0 else None}
                            for k, v in top[:12]}
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"breakdown_{wl.name}.json"), "w") as f:
            json.dump({"ms_timed_region": ms, "ms_profiled_pass": ms_prof, "steps": args.steps, "labels": breakdown}, f, indent=1)
    except OSError:
        pass
    if e2e is not None:
        out["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(wl, stats)
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
