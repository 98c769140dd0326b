"""ctypes loader for oracle/lsap.c (TEST INFRASTRUCTURE; see lsap.c header)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib():
    if not os.path.exists(_SO):
        build()
    lib = ctypes.CDLL(_SO)
    lib.ssb_oracle_lsap.restype = ctypes.c_int
    return lib


def linear_sum_assignment_c(cost):
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    nr, nc = cost.shape
    k = min(nr, nc)
    rows = np.zeros(k, dtype=np.int64)
    cols = np.zeros(k, dtype=np.int64)
    rc = _lib().ssb_oracle_lsap(cost.ctypes.data_as(ctypes.c_void_p), nr, nc,
                                rows.ctypes.data_as(ctypes.c_void_p),
                                cols.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError("cost matrix is infeasible")
    return rows, cols
