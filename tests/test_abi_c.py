"""The C ABI from plain C: include/tdvp_b200.h compiles as C11, every declared entry point links against
libtdvp_b200.so (CPU), and the same binary runs a ZGEMM on the GPU without Python in the loop (GPU)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(tmp_path):
    lib = os.path.join(ROOT, "pytdscf_b200", "libtdvp_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__

        __graft_entry__.build()
    exe = os.path.join(tmp_path, "abi_smoke")
    cmd = ["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-L", os.path.join(ROOT, "pytdscf_b200"), "-ltdvp_b200",
           "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm", "-Wl,-rpath," + os.path.join(ROOT, "pytdscf_b200"),
           "-Wl,-rpath," + os.path.join(CUDA, "lib64"), "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_header_is_c_and_all_entry_points_link(tmp_path):
    exe = _build(str(tmp_path))
    out = subprocess.run([exe, "link-only"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    from pytdscf_b200._lib import SIGNATURES

    assert f"entry_points={len(SIGNATURES)}" in out.stdout   # every symbol the header declares


@pytest.mark.gpu
def test_c_program_runs_a_gemm_on_the_gpu(tmp_path):
    exe = _build(str(tmp_path))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bad_arg_rc=-" in out.stdout
