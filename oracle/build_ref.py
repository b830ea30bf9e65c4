"""Build ``oracle/_ref``: the UNMODIFIED reference (PyTDSCF 1.3.3) byte-compiled for the GPU box -- TEST INFRASTRUCTURE.

    python oracle/build_ref.py            # build container only (needs /root/reference)

`/root/reference` does not exist on the GPU box, and reference SOURCES must not be copied into this repository.  What
travels instead is a build product, exactly like a compiled ``.so`` would for a C reference: every module under
``/root/reference/pytdscf`` is compiled from the sources WHERE THEY LIE to source-less bytecode, packed into ONE archive
``oracle/_ref/pytdscf_ref.zip`` (git-ignored, not gpurun-ignored; loose ``*.pyc`` files do not survive the snapshot to the
GPU box).  CPython's zipimport loads such an archive like the package itself, so
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
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/pytdscf"
OUT_DIR = os.path.join(HERE, "_ref")
DST = os.path.join(OUT_DIR, "pytdscf_ref.zip")


def build(verbose: bool = True) -> str | None:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"[build_ref] {SRC} not present: keeping the prebuilt oracle/_ref (if any)")
        return DST if os.path.exists(DST) else None
    if os.path.isdir(OUT_DIR):
        shutil.rmtree(OUT_DIR)
    os.makedirs(OUT_DIR)
    n = 0
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(DST, "w", zipfile.ZIP_DEFLATED) as zf:
        for dirpath, dirnames, filenames in os.walk(SRC):
            dirnames[:] = [d for d in dirnames if d != "__pycache__"]
            rel = os.path.relpath(dirpath, SRC)
            for fn in filenames:
                if not fn.endswith(".py"):
                    continue
                arc = os.path.normpath(os.path.join("pytdscf", rel, fn + "c"))   # legacy layout: module.pyc in place of module.py
                out = os.path.join(tmp, f"m{n}.pyc")
                py_compile.compile(os.path.join(dirpath, fn), cfile=out, dfile=os.path.join("pytdscf", rel, fn), doraise=True,
                                   optimize=0)
                zf.write(out, arc)
                n += 1
    meta = {"source": SRC, "modules": n, "python": sys.version.split()[0], "magic": importlib.util.MAGIC_NUMBER.hex(),
            "version": "1.3.3", "archive": os.path.basename(DST)}
    with open(os.path.join(OUT_DIR, "BUILD_INFO.json"), "w") as f:
        json.dump(meta, f, indent=1)
    if verbose:
        print(f"[build_ref] {n} modules byte-compiled into {DST}")
    return DST


if __name__ == "__main__":
    build()
