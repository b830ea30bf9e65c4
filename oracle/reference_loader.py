"""Import the UNMODIFIED reference (PyTDSCF 1.3.3, NumPy backend) from /root/reference.

TEST INFRASTRUCTURE ONLY (see oracle/refshim/README.md).  Used by tests/golden/make_golden.py in
the build container to pin the oracle; `/root/reference` does not exist on the GPU box, so nothing
in `-m gpu` tests, `smoke()` or `bench.py` calls this.
"""
from __future__ import annotations

import os
import sys

REFERENCE_ROOT = "/root/reference"
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pytdscf"))


def load_reference():
    """Return the imported ``pytdscf`` reference package (NumPy backend only)."""
    if not reference_available():
        raise RuntimeError(f"{REFERENCE_ROOT} is not present (only available in the build container)")
    sys.dont_write_bytecode = True
    for p in (REFERENCE_ROOT, _SHIM):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    os.environ.setdefault("LOGURU_LEVEL", "ERROR")
    import pytdscf  # noqa: PLC0415

    return pytdscf
