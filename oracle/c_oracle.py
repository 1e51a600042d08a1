"""ctypes wrapper of oracle/prophet_oracle.c (test / CPU-baseline infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_oracle

_lib = None


class COpts(C.Structure):
    _fields_ = [("growth", C.c_int), ("multiplicative", C.c_int), ("n_changepoints", C.c_int),
                ("changepoint_range", C.c_double), ("tau", C.c_double), ("seas_prior", C.c_double),
                ("yearly", C.c_int), ("weekly", C.c_int), ("daily", C.c_int), ("max_iter", C.c_int),
                ("init_alpha", C.c_double), ("tol_obj", C.c_double), ("tol_rel_obj", C.c_double),
                ("tol_grad", C.c_double), ("tol_rel_grad", C.c_double), ("tol_param", C.c_double),
                ("algorithm", C.c_int), ("reserved", C.c_int)]

ALG_LBFGS_NEWTON, ALG_LBFGS, ALG_NEWTON = 0, 1, 2


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle.build())
        _lib.po_fit_batch.restype = C.c_int
        _lib.po_fit_batch_trace.restype = C.c_int
        _lib.po_objective.restype = C.c_int
    return _lib


def options(growth="logistic", seasonality_mode="multiplicative", yearly=-1, weekly=-1, daily=-1) -> COpts:
    o = COpts()
    load().po_default_options(C.byref(o))
    o.growth = 1 if growth == "logistic" else 0
    o.multiplicative = 1 if seasonality_mode == "multiplicative" else 0
    o.yearly, o.weekly, o.daily = yearly, weekly, daily
    return o


def fit_batch(ds, y, offsets, floor=0.0, cap_multiplier=1.1, opts: COpts | None = None, nthreads: int = 0, pstride: int = 96,
              trace_cap: int = 0):
    """Returns (theta, f, info) -- and a 4th array [n, trace_cap, 4] of (iteration, f_k, alpha_k, n_evals) rows per
    accepted L-BFGS iteration when ``trace_cap`` > 0."""
    opts = opts or options()
    ds = np.ascontiguousarray(ds, dtype=np.int64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = offsets.size - 1
    theta = np.zeros((n, pstride))
    f = np.zeros(n)
    info = np.zeros((n, 5), np.int32)
    trace = np.zeros((n, trace_cap, 4)) if trace_cap > 0 else None
    load().po_fit_batch_trace(C.c_void_p(ds.ctypes.data), C.c_void_p(y.ctypes.data), C.c_void_p(offsets.ctypes.data), C.c_int(n),
                              C.c_double(floor), C.c_double(cap_multiplier), C.byref(opts), C.c_void_p(theta.ctypes.data),
                              C.c_int(pstride), C.c_void_p(f.ctypes.data), C.c_void_p(info.ctypes.data), C.c_int(nthreads),
                              C.c_void_p(trace.ctypes.data if trace is not None else None), C.c_int(trace_cap))
    if trace is not None:
        return theta, f, info, trace
    return theta, f, info     # info columns: status, iters, n_evals, S, K


def objective(ds, y, floor, cap, theta, opts: COpts | None = None):
    opts = opts or options()
    ds = np.ascontiguousarray(ds, dtype=np.int64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    g = np.zeros(96)
    f = C.c_double()
    S, K = C.c_int(), C.c_int()
    err = load().po_objective(C.c_void_p(ds.ctypes.data), C.c_void_p(y.ctypes.data), C.c_int(ds.size), C.c_double(floor),
                              C.c_double(cap), C.byref(opts), C.c_void_p(theta.ctypes.data), C.byref(f),
                              C.c_void_p(g.ctypes.data), C.byref(S), C.byref(K))
    return err, f.value, g[:S.value + K.value + 3]
