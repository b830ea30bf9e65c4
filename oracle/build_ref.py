"""Build ``oracle/_ref``: the UNMODIFIED reference (PyTDSCF 1.3.3) byte-compiled for the GPU box -- TEST INFRASTRUCTURE.

    python oracle/build_ref.py            # build container only (needs /root/reference)

`/root/reference` does not exist on the GPU box, and reference SOURCES must not be copied into this repository.  What
travels instead is a build product, exactly like a compiled ``.so`` would for a C reference: every module under
``/root/reference/pytdscf`` is compiled from the sources WHERE THEY LIE to a source-less ``.pyc`` under
``oracle/_ref/pytdscf`` (git-ignored, not gpurun-ignored).  CPython imports such a tree like the package itself, so
``bench.py --impl reference`` runs the reference's own ``MPSCoefMPO.propagate`` / ``propagate_along_sweep`` on the box's
host cores (``cpu_baseline.kind == "reference"``) under the import stubs of ``oracle/refshim``.  Potential-energy data
files (5 MB of .db / .npy) are not needed by the hot path and are left out.  Same interpreter (image) on both sides, so
the bytecode magic matches; ``oracle/reference_loader.py`` checks it and falls back to the oracle port otherwise."""
from __future__ import annotations

import importlib.util
import json
import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/pytdscf"
DST = os.path.join(HERE, "_ref", "pytdscf")


def build(verbose: bool = True) -> str | None:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"[build_ref] {SRC} not present: keeping the prebuilt oracle/_ref (if any)")
        return DST if os.path.isdir(DST) else None
    if os.path.isdir(os.path.dirname(DST)):
        shutil.rmtree(os.path.dirname(DST))
    n = 0
    for dirpath, dirnames, filenames in os.walk(SRC):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, SRC)
        for fn in filenames:
            if not fn.endswith(".py"):
                continue
            out = os.path.join(DST, rel, fn + "c")            # legacy layout: module.pyc next to where module.py would be
            os.makedirs(os.path.dirname(out), exist_ok=True)
            py_compile.compile(os.path.join(dirpath, fn), cfile=out, dfile=os.path.join("pytdscf", rel, fn), doraise=True,
                               optimize=0)
            n += 1
    meta = {"source": SRC, "modules": n, "python": sys.version.split()[0], "magic": importlib.util.MAGIC_NUMBER.hex(),
            "version": "1.3.3"}
    with open(os.path.join(os.path.dirname(DST), "BUILD_INFO.json"), "w") as f:
        json.dump(meta, f, indent=1)
    if verbose:
        print(f"[build_ref] {n} modules byte-compiled into {DST}")
    return DST


if __name__ == "__main__":
    build()
