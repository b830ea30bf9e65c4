"""Probe (GPU box): time the K_eff / H_eff GEMM shapes of config 2 (D = 64) under every tile configuration, and the
complete keff_apply / heff_apply calls, to find launch-path regressions.  CUDA events, 200 launches each."""
import sys

import torch

sys.path.insert(0, ".")
from pytdscf_b200._engine import Engine  # noqa: E402

eng = Engine(0)


def rnd(*s):
    return torch.randn(s, dtype=torch.complex128, device="cuda")


def timeit(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


for (M, N, K, ta, tb, label) in [(192, 64, 64, 0, 0, "keff.g1 w=3"), (128, 64, 64, 0, 0, "keff.g1 w=2"), (64, 64, 192, 0, 1, "keff.g2 w=3"),
                                 (192, 512, 64, 0, 0, "heff.s1 w=3"), (512, 64, 192, 0, 1, "heff.s3 w=3")]:
    A = rnd(*((K, M) if ta else (M, K)))
    B = rnd(*((N, K) if tb else (K, N)))
    C = rnd(M, N)
    row = []
    for cfg in ("auto", "tiny", "small", "big", "tma"):
        eng.set_gemm_config(cfg, 0, 0)
        row.append(f"{cfg} {timeit(lambda: eng.zgemm(A, B, ta, tb, 1.0, 0.0, C)):7.2f} us")
    eng.set_gemm_config("auto", 0, 0)
    print(f"{label:14s} {M}x{N}x{K}: " + " | ".join(row), flush=True)

D, d = 64, 8
for w in (2, 3):
    L, R, sig, psi = rnd(D, w, D), rnd(D, w, D), rnd(D, D), rnd(D, d, D)
    core = eng.upload_core(rnd(w, d, d, w).cpu().numpy())
    print(f"w={w}: keff_apply {timeit(lambda: eng.keff_apply([(L, R, 1.0)], sig)):7.2f} us   heff_apply {timeit(lambda: eng.heff_apply([(L, core, R, 1.0)], psi)):7.2f} us", flush=True)

# host enqueue time vs. GPU time: is the stream starved by the host, or are the kernels themselves slow?
import time  # noqa: E402


def split(fn, n=300):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) * 1e6 / n, (t2 - t0) * 1e6 / n


w = 3
L, R, sig, psi = rnd(D, w, D), rnd(D, w, D), rnd(D, D), rnd(D, d, D)
core = eng.upload_core(rnd(w, d, d, w).cpu().numpy())
A, B, Cm = rnd(192, 64), rnd(64, 512), rnd(192, 512)
x = torch.zeros(1024, dtype=torch.complex128, device="cuda")
for label, fn in [("torch x.add_(1) (1 launch)", lambda: x.add_(1.0)),
                  ("zgemm 192x512x64 (1 launch)", lambda: eng.zgemm(A, B, 0, 0, 1.0, 0.0, Cm)),
                  ("keff_apply (2 launches)", lambda: eng.keff_apply([(L, R, 1.0)], sig)),
                  ("heff_apply (3 launches)", lambda: eng.heff_apply([(L, core, R, 1.0)], psi))]:
    enq, tot = split(fn)
    print(f"{label:32s} host enqueue {enq:6.2f} us/call   enqueue + drain {tot:6.2f} us/call", flush=True)
