"""Import the UNMODIFIED reference (PyTDSCF 1.3.3, NumPy backend).

TEST INFRASTRUCTURE ONLY (see oracle/refshim/README.md).  In the build container the package is imported from
/root/reference (tests/golden/make_golden*.py pin the oracle with it).  /root/reference does not exist on the GPU box;
there ``bench.py --impl reference`` / the ``cpu_baseline`` leg import the byte-compiled build product ``oracle/_ref``
(made by oracle/build_ref.py from the sources where they lie; no source travels).  Nothing in ``pytdscf_b200`` imports this.
"""
from __future__ import annotations

import os
import sys

REFERENCE_ROOT = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "refshim")
BUILT_DIR = os.path.join(_HERE, "_ref")
BUILT_ROOT = os.path.join(BUILT_DIR, "pytdscf_ref.zip")   # sys.path entry (zipimport)


def _built_ok() -> bool:
    import importlib.util
    import json

    info = os.path.join(BUILT_DIR, "BUILD_INFO.json")
    if not (os.path.exists(BUILT_ROOT) and os.path.exists(info)):
        return False
    try:
        return json.load(open(info)).get("magic") == importlib.util.MAGIC_NUMBER.hex()
    except Exception:
        return False


def reference_root(allow_built: bool = True) -> str | None:
    """Directory to put on sys.path: the source tree in the build container, else the byte-compiled oracle/_ref."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "pytdscf")):
        return REFERENCE_ROOT
    if allow_built and _built_ok():
        return BUILT_ROOT
    return None


def reference_available(allow_built: bool = False) -> bool:
    return reference_root(allow_built) is not None


def load_reference(allow_built: bool = False, prefer_built: bool = False):
    """Return the imported ``pytdscf`` reference package (NumPy backend only).  ``prefer_built`` imports oracle/_ref even
    when /root/reference exists (used to test the build product in the build container)."""
    root = BUILT_ROOT if (prefer_built and _built_ok()) else reference_root(allow_built)
    if root is None:
        raise RuntimeError(f"{REFERENCE_ROOT} is not present (only available in the build container) and oracle/_ref "
                           "was not built (python oracle/build_ref.py)")
    sys.dont_write_bytecode = True
    for p in (root, _SHIM):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    os.environ.setdefault("LOGURU_LEVEL", "ERROR")
    import pytdscf  # noqa: PLC0415

    return pytdscf
