"""Wire format of the ``model`` binary column of the models table.

The reference stores ``pickle.dumps(Prophet object)`` there (src/jobs/prophet_modeler.py:72-73,
read back by src/jobs/prophet_scorer.py:48): tens of KB per series carrying the whole
history frame.  The batched scorer needs only the fitted parameters and scaling metadata, so
the record is a fixed-layout little-endian struct (about 0.7 KB):

    magic 'PB2M' | u16 version | u16 flags (1 logistic, 2 multiplicative) | i32 smax | i32 kmax
    | i32[4] option switches (yearly, weekly, daily as PB200_SEAS_AUTO/0/1, n_changepoints)
    | i32[8] meta_i32 (T, S, n_changepoints_real, seasonality mask, status, iters, n_evals, i1)
    | i64[2] meta_i64 (start_ns, t_scale_ns) | i64 last_ds_ns (history_dates.max())
    | f64[4] meta_f64 (y_scale, floor, cap, neg_log_posterior)
    | f64[pstride] params (k, m, sigma_obs, delta[smax], beta[kmax]) | f64[smax] t_change

Encoding / decoding is vectorised over the whole batch (one numpy structured array),
never a per-row Python loop.
"""
from __future__ import annotations

import numpy as np
import pyarrow as pa

from .batched import FittedBatch

MAGIC = b"PB2M"
VERSION = 1
FLAG_LOGISTIC, FLAG_MULT = 1, 2


def record_dtype(smax: int, kmax: int) -> np.dtype:
    pstride = 3 + smax + kmax
    return np.dtype([("magic", "S4"), ("version", "<u2"), ("flags", "<u2"), ("smax", "<i4"), ("kmax", "<i4"),
                     ("switches", "<i4", (4,)), ("meta_i32", "<i4", (8,)), ("meta_i64", "<i8", (2,)), ("last_ds", "<i8"),
                     ("meta_f64", "<f8", (4,)), ("params", "<f8", (pstride,)), ("tchange", "<f8", (smax,))])


def encode(fitted: FittedBatch, last_ds_ns: np.ndarray, opts) -> pa.Array:
    """FittedBatch (+ the pb200 Options it was fitted with) -> Arrow binary array, one record per model."""
    logistic, multiplicative = opts.growth == 1, bool(opts.multiplicative)
    fitted = fitted.to_host()
    n = fitted.n
    dt = record_dtype(fitted.smax, fitted.kmax)
    rec = np.zeros(n, dtype=dt)
    rec["magic"] = MAGIC
    rec["version"] = VERSION
    rec["flags"] = (FLAG_LOGISTIC if logistic else 0) | (FLAG_MULT if multiplicative else 0)
    rec["smax"], rec["kmax"] = fitted.smax, fitted.kmax
    rec["switches"] = np.array([opts.yearly, opts.weekly, opts.daily, opts.n_changepoints], dtype=np.int32)
    rec["meta_i32"], rec["meta_i64"], rec["meta_f64"] = fitted.meta_i32, fitted.meta_i64, fitted.meta_f64
    rec["last_ds"] = np.asarray(last_ds_ns, dtype=np.int64)
    rec["params"], rec["tchange"] = fitted.params, fitted.tchange
    size = dt.itemsize
    offsets = pa.py_buffer((np.arange(n + 1, dtype=np.int64) * size).astype(np.int32).tobytes()) \
        if n * size < 2**31 else None
    data = pa.py_buffer(rec.tobytes())
    if offsets is not None:
        return pa.Array.from_buffers(pa.binary(), n, [None, offsets, data])
    off64 = pa.py_buffer((np.arange(n + 1, dtype=np.int64) * size).tobytes())
    return pa.Array.from_buffers(pa.large_binary(), n, [None, off64, data])


def decode(col) -> tuple:
    """Arrow binary column -> (FittedBatch, last_ds_ns, dict of the fit-time options)."""
    if isinstance(col, pa.ChunkedArray):
        col = col.combine_chunks() if col.num_chunks != 1 else col.chunk(0)
    n = len(col)
    if n == 0:
        raise ValueError("empty model column")
    if col.null_count:
        raise ValueError("model column holds nulls (the reference returns an empty frame for those rows)")
    first = col[0].as_py()
    if first[:4] != MAGIC:
        raise ValueError("model blob is not a PB2M record (fbprophet pickles cannot be scored on this path)")
    smax, kmax = np.frombuffer(first[8:16], dtype="<i4")
    dt = record_dtype(int(smax), int(kmax))
    bufs = col.buffers()
    off_dt = np.int64 if pa.types.is_large_binary(col.type) else np.int32
    offs = np.frombuffer(bufs[1], dtype=off_dt)[col.offset:col.offset + n + 1]
    if not np.all(np.diff(offs) == dt.itemsize):
        raise ValueError("model records of differing layout in one table")
    rec = np.frombuffer(bufs[2], dtype=dt, count=n, offset=int(offs[0]))
    if not (np.all(rec["magic"] == MAGIC) and np.all(rec["version"] == VERSION)):
        raise ValueError("bad model record header")
    flags = int(rec["flags"][0])
    fb = FittedBatch(np.ascontiguousarray(rec["params"]), np.ascontiguousarray(rec["tchange"]),
                     np.ascontiguousarray(rec["meta_i32"]), np.ascontiguousarray(rec["meta_i64"]),
                     np.ascontiguousarray(rec["meta_f64"]), int(smax), int(kmax))
    if not np.all(rec["switches"] == rec["switches"][0]) or not np.all(rec["flags"] == flags):
        raise ValueError("model records fitted with differing options in one table")
    sw = rec["switches"][0]
    info = {"logistic": bool(flags & FLAG_LOGISTIC), "multiplicative": bool(flags & FLAG_MULT),
            "yearly": int(sw[0]), "weekly": int(sw[1]), "daily": int(sw[2]), "n_changepoints": int(sw[3])}
    return fb, np.ascontiguousarray(rec["last_ds"]), info


def from_fbprophet_pickle(blobs, floor=None, cap=None):
    """Models written by the REFERENCE (``pickle.dumps(Prophet object)``, src/jobs/prophet_modeler.py:72-73) ->
    (FittedBatch, last_ds_ns, options dict), so that genuine fbprophet fits can be scored on the GPU path -- and
    compared with this repo's fits of the same input, the strongest parity check there is (BASELINE.md
    section 2).

    Unpickling a Prophet object needs ``fbprophet`` (0.5, the reference's pin) or ``prophet`` importable; neither
    exists in this image (no network), so this function has NOT been exercised against a real pickle: it follows
    fbprophet 0.5's attribute names as recalled -- ``params`` {'k','m','delta','sigma_obs','beta'} (1 x n arrays),
    ``changepoints_t``, ``start``, ``t_scale``, ``y_scale``, ``seasonalities`` (OrderedDict name -> period /
    fourier_order / mode), ``growth``, ``seasonality_mode``, ``history_dates`` -- and raises ImportError with the
    reason when the class cannot be imported.  Only the default seasonalities (yearly 10, weekly 3, daily 4) map
    onto the compiled kernels; anything else is refused."""
    import pickle
    try:
        try:
            import fbprophet  # noqa: F401
        except ImportError:
            import prophet  # noqa: F401
    except ImportError as exc:
        raise ImportError("reading the reference's model pickles needs fbprophet (or prophet) importable: "
                          "pickle.loads re-creates a fbprophet.forecaster.Prophet object") from exc
    models = [pickle.loads(b) for b in blobs]
    n = len(models)
    if n == 0:
        raise ValueError("no models")
    m0 = models[0]
    logistic = m0.growth == "logistic"
    mult = getattr(m0, "seasonality_mode", "additive") == "multiplicative"
    orders = {"yearly": 10, "weekly": 3, "daily": 4}
    smax = max(1, max(len(np.atleast_1d(m.changepoints_t)) for m in models))
    sw = {k: 0 for k in orders}
    for m in models:
        for name, spec in m.seasonalities.items():
            if name not in orders or int(spec["fourier_order"]) != orders[name]:
                raise ValueError(f"seasonality {name!r} (order {spec.get('fourier_order')}) has no compiled kernel")
            sw[name] = 1
    kmax = sum(2 * orders[k] for k in orders if sw[k]) or 1
    pstride = 3 + smax + kmax
    params = np.zeros((n, pstride))
    tchange = np.zeros((n, smax))
    mi32 = np.zeros((n, 8), np.int32)
    mi64 = np.zeros((n, 2), np.int64)
    mf64 = np.zeros((n, 4))
    last = np.zeros(n, np.int64)
    for i, m in enumerate(models):
        p = {k: np.asarray(v, dtype=np.float64).reshape(-1) for k, v in m.params.items()}
        cps = np.atleast_1d(np.asarray(m.changepoints_t, dtype=np.float64))
        S = len(cps)
        mask, col = 0, 0
        beta = p["beta"]
        for bit, name in ((1, "yearly"), (2, "weekly"), (4, "daily")):
            if name in m.seasonalities:
                mask |= bit
        # fbprophet orders the seasonal columns as the seasonalities were added (yearly, weekly, daily for the defaults)
        k = 0
        for bit, name in ((1, "yearly"), (2, "weekly"), (4, "daily")):
            if sw[name]:
                if mask & bit:
                    params[i, 3 + smax + col:3 + smax + col + 2 * orders[name]] = beta[k:k + 2 * orders[name]]
                    k += 2 * orders[name]
                col += 2 * orders[name]
        params[i, 0], params[i, 1], params[i, 2] = p["k"][0], p["m"][0], p["sigma_obs"][0]
        params[i, 3:3 + S] = p["delta"][:S]
        tchange[i, :S] = cps
        start = np.datetime64(m.start, "ns").astype(np.int64)
        t_scale = np.timedelta64(m.t_scale, "ns").astype(np.int64)
        hist_max = np.datetime64(max(m.history_dates), "ns").astype(np.int64)
        mi32[i] = (len(m.history_dates), S, S, mask, 31, 0, 0, 0)
        mi64[i] = (start, t_scale)
        mf64[i] = (float(m.y_scale), 0.0 if floor is None else float(np.atleast_1d(floor)[min(i, np.size(floor) - 1)]),
                   np.nan if cap is None else float(np.atleast_1d(cap)[min(i, np.size(cap) - 1)]), np.nan)
        last[i] = hist_max
    fb = FittedBatch(params, tchange, mi32, mi64, mf64, smax, kmax)
    info = {"logistic": logistic, "multiplicative": mult, "yearly": sw["yearly"] or 0, "weekly": sw["weekly"] or 0,
            "daily": sw["daily"] or 0, "n_changepoints": smax}
    return fb, last, info
