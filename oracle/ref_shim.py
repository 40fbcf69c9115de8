"""Import the real reference (read-only tree at /root/reference) in the BUILD container.

Used only by tests/golden/make_golden.py (fixture generation) and by optional live
cross-checks that skip when /root/reference is absent (it does not exist on the GPU box).
pynndescent/__init__.py:22 calls importlib.metadata.version("pynndescent"), which fails
for an un-installed tree, hence the two-line shim.
"""
import importlib.metadata as _im
import os
import sys

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pynndescent"))


def import_reference():
    if not available():
        raise ImportError("reference tree not present")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    _orig = _im.version
    _im.version = lambda n: "0.6.0" if n == "pynndescent" else _orig(n)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import pynndescent  # noqa: F401
    return pynndescent
