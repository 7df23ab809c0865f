"""Python driver for the CPU oracle (TEST INFRASTRUCTURE - see ugf_oracle.cpp).

Binds oracle/_build/libugf_oracle.so (prefix ugfo_) to the same host-side cloud class
the product uses, so that a test can run one case through both and compare.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
import os
import subprocess

from unigasfoam_b200 import _capi
from unigasfoam_b200.cloud import UniGasCloud

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libugf_oracle.so")
_api = None


def build(force=False):
    src = os.path.join(_HERE, "ugf_oracle.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB


def api():
    global _api
    if _api is None:
        if not os.path.exists(LIB):
            build()
        _api = _capi.Api(LIB, "ugfo_")
    return _api


class OracleCloud(UniGasCloud):
    migrate_buffers = "host"

    def __init__(self, *a, **kw):
        kw["api"] = api()
        super().__init__(*a, **kw)


def num_threads():
    import ctypes
    f = api().lib.ugfo_num_threads
    f.restype = ctypes.c_int
    return f()


def set_num_threads(n):
    """OpenMP threads of the oracle's parallel loops (bench.py: all host cores, whatever OMP_NUM_THREADS the launcher left)."""
    import ctypes
    f = api().lib.ugfo_set_num_threads
    f.argtypes = [ctypes.c_int]
    f.restype = None
    f(int(n))
    return num_threads()
