"""Time the DMMA ZGEMM (tdvp_zgemm) on the GEMM shapes the TDVP workloads issue; CUDA events, best of 5."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pytdscf_b200._engine import Engine  # noqa: E402

SHAPES = [
    # (label, M, N, K, transA, transB)
    ("c4 heff stage1  NN", 8192, 4096, 1024, 0, 0),
    ("c4 heff stage3  NT", 4096, 1024, 8192, 0, 1),
    ("c4 env  step3   CN", 1024, 8192, 4096, 2, 0),
    ("c4 heff stage2  NN (d=16,w=8)", 1048576, 128, 128, 0, 0),
    ("c4 heff stage2  NN (d=4,w=8)", 1048576, 32, 32, 0, 0),
    ("c4 keff gemm1   NN", 8192, 1024, 1024, 0, 0),
    ("c4 keff gemm2   NT", 1024, 1024, 8192, 0, 1),
    ("c4e heff stage1 NN", 8192, 16384, 1024, 0, 0),
    ("c3 heff stage1  NN", 1280, 2560, 256, 0, 0),
    ("c3 heff stage3  NT", 2560, 256, 1280, 0, 1),
    ("c3 env  step3   CN", 256, 1280, 2560, 2, 0),
    ("c3 keff gemm1   NN", 1280, 256, 256, 0, 0),
    ("c3 keff gemm2   NT", 256, 256, 1280, 0, 1),
    ("c3s heff stage1 NN (id skip)", 1024, 2560, 256, 0, 0),
    ("c3s heff stage3 NT (id skip)", 2560, 256, 1024, 0, 1),
    ("c5 heff stage1  NN pot", 1536, 4096, 512, 0, 0),
    ("c5 heff stage1  NN kin", 512, 4096, 512, 0, 0),
    ("c5 heff stage3  NT pot", 4096, 512, 1536, 0, 1),
    ("c5 heff stage3  NT kin", 4096, 512, 512, 0, 1),
    ("c5 keff gemm1   NN", 1536, 512, 512, 0, 0),
    ("c5 keff gemm2   NT", 512, 512, 1536, 0, 1),
    ("square 4096     NN", 4096, 4096, 4096, 0, 0),
    ("square 8192     NN", 8192, 8192, 8192, 0, 0),
]


def main():
    """usage: zgemm_shapes.py [label-filter|-] [auto|big|small|tiny|tma]"""
    eng = Engine(0)
    out = []
    only = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] != "-" else None
    cfg = sys.argv[2] if len(sys.argv) > 2 else "auto"
    eng.set_gemm_config(cfg, 0, 0)
    for label, M, N, K, ta, tb in SHAPES:
        if only and only not in label:
            continue
        A = torch.randn((K, M) if ta else (M, K), dtype=torch.complex128, device="cuda")
        B = torch.randn((N, K) if tb else (K, N), dtype=torch.complex128, device="cuda")
        C = torch.empty((M, N), dtype=torch.complex128, device="cuda")
        for _ in range(2):
            eng.zgemm(A, B, ta, tb, 1.0, 0.0, C)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.zgemm(A, B, ta, tb, 1.0, 0.0, C)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        tf = 8.0 * M * N * K / (best * 1e-3) / 1e12
        opA = {0: A, 1: A.T, 2: A.conj().T}[ta]
        opB = {0: B, 1: B.T, 2: B.conj().T}[tb]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.matmul(opA, opB)
        t0.record()
        ref = torch.matmul(opA, opB)
        t1.record()
        t1.synchronize()
        tf_cublas = 8.0 * M * N * K / (t0.elapsed_time(t1) * 1e-3) / 1e12
        err = float((C - ref).abs().max() / ref.abs().max())
        rec = {"cfg": cfg, "shape": label, "M": M, "N": N, "K": K, "ms": round(best, 4), "tflops": round(tf, 2),
               "frac_of_37": round(tf / 37.0, 3), "cublas_tflops": round(tf_cublas, 2), "relerr": err}
        print(json.dumps(rec))
        out.append(rec)
    eng.close()


if __name__ == "__main__":
    main()
