"""ctypes wrapper of the plain-C oracle (oracle/nabo_oracle.c).  TEST / BASELINE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
_L = None


def build() -> str:
    src = os.path.join(HERE, "nabo_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "_build/liboracle.so"])
    return LIB


def lib():
    global _L
    if _L is None:
        build()
        _L = C.CDLL(LIB)
    return _L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def dist(x, y, metric: str, f: float = 0.25):
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    d = np.empty((x.shape[0], y.shape[0]))
    if metric == "euclidean":
        lib().oracle_euclidean_dist(_p(x), _p(y), _p(d), x.shape[0], y.shape[0], x.shape[1])
    else:
        lib().oracle_mod_canberra_dist(_p(x), _p(y), _p(d), x.shape[0], y.shape[0], x.shape[1], C.c_double(f))
    return d


def _slices(n, parts):
    b = [n * i // parts for i in range(parts + 1)]
    return [(b[i], b[i + 1]) for i in range(parts) if b[i + 1] > b[i]]


def _run(fn, n, nthreads):
    """Run fn(lo, hi) over disjoint row slices; ctypes releases the GIL during the C call."""
    sl = _slices(n, max(1, int(nthreads)))
    if len(sl) <= 1:
        for lo, hi in sl:
            fn(lo, hi)
        return
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(len(sl)) as ex:
        list(ex.map(lambda s: fn(*s), sl))


def knn(q, r, k, metric="euclidean", f=0.25, mask=None, drop_first=False, nthreads=1):
    q = np.ascontiguousarray(q, np.float64)
    r = np.ascontiguousarray(r, np.float64)
    idx = np.empty((q.shape[0], k), np.int32)
    dst = np.empty((q.shape[0], k), np.float64)
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    L = lib()

    def work(lo, hi):
        L.oracle_knn(_p(q[lo:hi]), _p(r), hi - lo, r.shape[0], q.shape[1], int(k),
                     0 if metric == "euclidean" else 1, C.c_double(f), None if m is None else _p(m),
                     int(bool(drop_first)), _p(idx[lo:hi]), _p(dst[lo:hi]))
    _run(work, q.shape[0], nthreads)
    return idx, dst


def snn_counts(tgt_knn, ref_knn, nthreads=1):
    t = np.ascontiguousarray(tgt_knn, np.int32)
    r = np.ascontiguousarray(ref_knn, np.int32)
    out = np.empty(t.shape, np.uint8)
    L = lib()
    _run(lambda lo, hi: L.oracle_snn_counts(_p(t[lo:hi]), _p(r), hi - lo, t.shape[1], _p(out[lo:hi])),
         t.shape[0], nthreads)
    return out


def scores(tgt_knn, counts, lut, n_ref, mult=1000.0):
    t = np.ascontiguousarray(tgt_knn, np.int32)
    c = np.ascontiguousarray(counts, np.uint8)
    lut = np.ascontiguousarray(lut, np.float64)
    out = np.empty(n_ref, np.float64)
    lib().oracle_scores(_p(t), _p(c), _p(lut), t.shape[0], t.shape[1], int(n_ref), C.c_double(mult), _p(out))
    return out
