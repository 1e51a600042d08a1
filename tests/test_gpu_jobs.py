"""GPU end-to-end tests of the drop-in jobs: the reference's own test flow
(tests/unit/prophet_modeler_test.py:59-75, tests/unit/prophet_scorer_test.py:83-114) with Spark
replaced by the batched operators, on the reference's fixture (config #1)."""
import os

import numpy as np
import pyarrow.dataset as pads
import pytest

from time_series_spark_b200.jobs.prophet_modeler import ProphetModeler, model_time_series
from time_series_spark_b200.jobs.prophet_scorer import ProphetScorer, forecast_time_series

pytestmark = pytest.mark.gpu


def test_model_then_forecast_time_series(tmp_path, model_input_dir, golden_oracle):
    mconfig = {"io": {"input": model_input_dir, "models": str(tmp_path / "build" / "models")},
               "model": {"floor": 0, "cap_multiplier": 1.1}}
    modeler = ProphetModeler(mconfig)
    spark_input_df = modeler.read_input_dataframe(None)
    # ---- test_model_time_series ----
    output_df = spark_input_df.groupby("series_id", "dim_id").apply(model_time_series(modeler.config))
    assert output_df.count() == 2
    assert output_df.columns == ["series_id", "dim_id", "floor", "cap", "model"]
    assert output_df.filter("series_id = 751 and dim_id = 91").count() == 1
    assert output_df.filter("series_id = 751 and dim_id = 155").count() == 1
    caps = dict(zip(output_df.table["dim_id"].to_pylist(), output_df.table["cap"].to_pylist()))
    assert caps[91] == np.float32(103591.40000000001) and caps[155] == np.float32(140054.2)   # FloatType column
    modeler.persist_models(output_df)
    model_df = pads.dataset(mconfig["io"]["models"], format="parquet").to_table()
    assert model_df.num_rows == 2
    assert model_df.column_names == ["series_id", "dim_id", "floor", "cap", "model"]

    # ---- test_read_model_dataframe / test_forecast_time_series ----
    sconfig = {"io": {"models": mconfig["io"]["models"], "forecasts": str(tmp_path / "build" / "forecasts")},
               "forecast": {"periods": 40, "frequency": "15min"}}
    scorer = ProphetScorer(sconfig)
    spark_model_df = scorer.read_model_dataframe(None)
    assert spark_model_df.columns == ["series_id", "dim_id", "floor", "cap", "model"]
    assert spark_model_df.select("series_id").distinct().count() == 1
    assert spark_model_df.select("dim_id").distinct().count() == 2
    assert spark_model_df.count() == 2
    output_df = spark_model_df.groupby("series_id", "dim_id").apply(forecast_time_series(scorer.config))
    assert output_df.count() == 80
    assert output_df.columns == ["series_id", "dim_id", "ds", "yhat"]
    assert output_df.filter("series_id = 751 and dim_id = 91").count() == 40
    assert output_df.filter("series_id = 751 and dim_id = 155").count() == 40
    # numerical contract the reference never asserted: forecast close to the oracle's golden vector
    for dim in (91, 155):
        sub = output_df.filter(f"series_id = 751 and dim_id = {dim}").table
        got = np.asarray(sub["yhat"].to_pylist(), dtype=np.float64)
        ts = np.asarray(sub["ds"].cast("int64").to_pylist(), dtype=np.int64)
        assert np.array_equal(ts, golden_oracle[f"d{dim}_future_ns"])
        ys = float(golden_oracle[f"d{dim}_y_scale"])
        assert np.max(np.abs(got - golden_oracle[f"d{dim}_yhat_int"])) <= 3e-2 * ys + 1
    converted_df = scorer.convert_forecasts(output_df)
    scorer.write_forecasts(converted_df)
    read_output_df = pads.dataset(sconfig["io"]["forecasts"], format="csv").to_table()
    assert read_output_df.column_names == ["created_timestamp", "series_id", "dim_id", "forecast_date",
                                           "forecast_timestamp", "forecast_quantity"]
    assert read_output_df.num_rows == 80


def test_static_entry_points_and_intervals(tmp_path, model_input_dir):
    mconfig = {"io": {"input": model_input_dir, "models": str(tmp_path / "models")},
               "model": {"floor": 0, "cap_multiplier": 1.1}}
    ProphetModeler.model(None, mconfig)
    sconfig = {"io": {"models": mconfig["io"]["models"], "forecasts": str(tmp_path / "forecasts")},
               "forecast": {"periods": 8, "frequency": "W", "intervals": True, "seed": 5}}
    ProphetScorer.score(None, sconfig)
    out = pads.dataset(sconfig["io"]["forecasts"], format="csv").to_table()
    assert out.num_rows == 16
    assert out.column_names[-2:] == ["yhat_lower", "yhat_upper"]
    lo, hi, q = (np.asarray(out[c].to_pylist(), dtype=float) for c in ("yhat_lower", "yhat_upper", "forecast_quantity"))
    assert np.all(lo < hi) and np.all(q >= 0)


def test_gpu_pack_matches_host_pack():
    """pack_groups_cuda (radix sorts on the GPU) == pack_groups (numpy lexsort), including null-y rows."""
    import pyarrow as pa
    import torch
    from time_series_spark_b200.pack import pack_groups, pack_groups_cuda
    rng = np.random.RandomState(3)
    n = 20000
    sid = rng.randint(0, 7, n).astype(np.int32)
    did = rng.randint(0, 50, n).astype(np.int32)
    ds = (rng.randint(0, 4000, n).astype(np.int64) * 900 * 10**9)
    y = rng.randint(1, 1000, n).astype(np.int32)
    mask = rng.rand(n) < 0.02
    tbl = pa.table({"series_id": pa.array(sid), "dim_id": pa.array(did),
                    "ds": pa.array(ds, pa.int64()).cast(pa.timestamp("ns")),
                    "y": pa.array(y, pa.int32(), mask=mask)})
    h = pack_groups(tbl, pin=False)
    g = pack_groups_cuda(tbl)
    assert g.on_device and not h.on_device
    assert np.array_equal(h.series_id, g.series_id) and np.array_equal(h.dim_id, g.dim_id)
    assert np.array_equal(h.offsets, g.offsets) and np.array_equal(h.last_ds, g.last_ds)
    assert np.array_equal(h.n_rows_in, g.n_rows_in)
    assert np.array_equal(h.ds, g.ds.cpu().numpy())
    # rows with equal (series, dim, ds) may be ordered differently by the two stable sorts only if the input
    # order differs, which it does not: y must match exactly as well
    assert np.array_equal(h.y, g.y.cpu().numpy())


def _same_pack(h, g):
    assert np.array_equal(h.series_id, g.series_id) and np.array_equal(h.dim_id, g.dim_id)
    assert np.array_equal(h.offsets, g.offsets) and np.array_equal(h.last_ds, g.last_ds)
    assert np.array_equal(h.n_rows_in, g.n_rows_in)
    assert np.array_equal(h.ds, g.ds.cpu().numpy()) and np.array_equal(h.y, g.y.cpu().numpy())
    assert h.y.dtype == g.y.cpu().numpy().dtype


@pytest.mark.parametrize("case", ["chunked_seconds_negative_ids", "already_sorted", "float_y_with_nan", "bigger_than_a_slot",
                                  "int64_columns"])
def test_gpu_pack_upload_paths(case):
    """The column upload of pack_groups_cuda (Arrow chunks -> pinned ring -> HBM, key and unit conversion on the device,
    sort skipped for ordered input) against the host pack on the shapes it has to survive."""
    import pyarrow as pa
    from time_series_spark_b200 import pack
    rng = np.random.RandomState(8)
    if case == "chunked_seconds_negative_ids":
        parts = []
        for c in range(7):                                        # 7 chunks of uneven size, one of them empty
            n = [0, 1, 513, 4000, 37, 2048, 900][c]
            parts.append(pa.table({"series_id": pa.array(rng.randint(-3, 4, n).astype(np.int32)),
                                   "dim_id": pa.array(rng.randint(-40, 40, n).astype(np.int32)),
                                   "ds": pa.array(rng.randint(0, 3000, n).astype(np.int64) * 900, pa.int64()).cast(pa.timestamp("s")),
                                   "y": pa.array(rng.randint(1, 99, n).astype(np.int32))}))
        tbl = pa.concat_tables(parts)
        assert tbl["ds"].num_chunks >= 6
    elif case == "already_sorted":
        n = 30000
        sid = np.repeat(np.arange(30, dtype=np.int32), 1000)
        did = np.tile(np.repeat(np.arange(10, dtype=np.int32), 100), 30)
        ds = np.tile(np.arange(100, dtype=np.int64) * 900 * 10**9, 300)
        tbl = pa.table({"series_id": sid, "dim_id": did, "ds": pa.array(ds).cast(pa.timestamp("ns")),
                        "y": pa.array(rng.randint(1, 99, n).astype(np.int32))})
    elif case == "float_y_with_nan":
        n = 5000
        y = rng.rand(n) * 100
        y[rng.rand(n) < 0.03] = np.nan
        tbl = pa.table({"series_id": pa.array(rng.randint(0, 3, n).astype(np.int32)), "dim_id": pa.array(rng.randint(0, 9, n).astype(np.int32)),
                        "ds": pa.array(rng.randint(0, 900, n).astype(np.int64) * 60 * 10**9).cast(pa.timestamp("ns")),
                        "y": pa.array(y, pa.float64(), mask=rng.rand(n) < 0.01)})
    elif case == "bigger_than_a_slot":
        old = pack._SLOT_BYTES
        pack._SLOT_BYTES, pack._slots = 4096, {}                   # 512 int64 per slot: every column wraps the ring many times
        try:
            n = 9001
            tbl = pa.table({"series_id": pa.array(rng.randint(0, 5, n).astype(np.int32)), "dim_id": pa.array(rng.randint(0, 50, n).astype(np.int32)),
                            "ds": pa.array(rng.randint(0, 4000, n).astype(np.int64) * 10**9).cast(pa.timestamp("ns")),
                            "y": pa.array(rng.randint(1, 999, n).astype(np.int32))})
            _same_pack(pack.pack_groups(tbl, pin=False), pack.pack_groups_cuda(tbl))
        finally:
            pack._SLOT_BYTES, pack._slots = old, {}
        return
    else:
        n = 4000
        tbl = pa.table({"series_id": pa.array(rng.randint(0, 5, n).astype(np.int64)), "dim_id": pa.array(rng.randint(0, 50, n).astype(np.int64)),
                        "ds": pa.array(rng.randint(0, 4000, n).astype(np.int64) * 10**9),             # plain int64 ns
                        "y": pa.array(rng.randint(1, 999, n).astype(np.int64))})
    _same_pack(pack.pack_groups(tbl, pin=False), pack.pack_groups_cuda(tbl))


def _arrow_csv_rows(created, sid, did, ds_ns, qty):
    import io
    import pyarrow as pa
    import pyarrow.compute as pc
    import pyarrow.csv as pacsv
    ts = pa.array(ds_ns).cast(pa.timestamp("ns"))
    ms = pc.cast(ts, pa.timestamp("ms"), safe=False)
    tbl = pa.table({"created_timestamp": pa.array([created] * len(sid)), "series_id": sid, "dim_id": did,
                    "forecast_date": pc.strftime(ts, format="%Y-%m-%d"),
                    "forecast_timestamp": pc.strftime(ms, format="%Y-%m-%dT%H:%M:%SZ"), "forecast_quantity": qty})
    buf = io.BytesIO()
    pacsv.write_csv(tbl, buf, write_options=pacsv.WriteOptions(include_header=False, quoting_style="needed"))
    return buf.getvalue()


def test_gpu_csv_rows_match_arrow_writer(gpu_ctx):
    """pb200_forecast_csv_{lengths,rows}_device against pyarrow's writer, byte for byte: extreme and negative integers,
    sub-millisecond timestamps, leap days, the last int64 nanosecond, a row count that is not a multiple of the warp."""
    import torch
    from time_series_spark_b200 import batched
    rng = np.random.RandomState(0)
    n = 100_003
    sid = rng.randint(-5, 2**31 - 1, n).astype(np.int32)
    did = rng.randint(-2**31, 2**31 - 1, n).astype(np.int32)
    qty = rng.randint(-2**31, 2**31 - 1, n).astype(np.int32)
    sid[:5] = [0, -1, 2**31 - 1, -2**31, 10]
    qty[:4] = [0, 9, 10, -10]
    ds = rng.randint(0, 9_223_372_036, n).astype(np.int64) * 10**9 + rng.randint(0, 10**9, n)
    ds[:6] = [0, 999_999, 86399_999_999_999, 951782400 * 10**9, 4107542400 * 10**9 - 1, 2**63 - 1]
    # a realistic block too: small ids, one horizon repeated
    sid[50_000:] = np.arange(n - 50_000) // 672 // 100
    did[50_000:] = np.arange(n - 50_000) // 672 % 100
    ds[50_000:] = 1_650_000_000 * 10**9 + 900 * 10**9 * (np.arange(n - 50_000) % 672)
    qty[50_000:] = rng.randint(0, 200_000, n - 50_000)
    created = "2026-09-23T04:10:26+00:00"
    want = _arrow_csv_rows(created, sid, did, ds, qty)
    for m in (n, 1, 31, 32, 33):
        got = batched.forecast_csv_device(gpu_ctx, torch.from_numpy(sid[:m].copy()).cuda(), torch.from_numpy(did[:m].copy()).cuda(),
                                          torch.from_numpy(ds[:m].copy()).cuda(), torch.from_numpy(qty[:m].copy()).cuda(),
                                          created.encode()).cpu().numpy().tobytes()
        ref = want if m == n else b"".join(l + b"\n" for l in want.split(b"\n")[:m])
        assert got == ref, (m, got[:200], ref[:200])
    assert batched.forecast_csv_row_host(3, -4, 951782400 * 10**9, 7, created.encode()) == \
        b'"2026-09-23T04:10:26+00:00",3,-4,"2000-02-29","2000-02-29T00:00:00.000Z",7\n'


def test_gpu_forecast_writer_equals_arrow_writer(tmp_path, monkeypatch):
    """write_forecasts with forecast.writer = gpu (part files formatted by the kernels) vs arrow: same rows, same bytes."""
    import pyarrow as pa
    from time_series_spark_b200.frame import Frame
    from time_series_spark_b200.jobs import prophet_scorer as ps
    rng = np.random.RandomState(5)
    n_models, H = 150, 96
    grid = 1_650_000_000 * 10**9 + 900 * 10**9 * np.arange(H, dtype=np.int64)
    tbl = pa.table({"series_id": pa.array(np.repeat(np.arange(n_models, dtype=np.int32) // 10 - 3, H)),
                    "dim_id": pa.array(np.repeat(np.arange(n_models, dtype=np.int32) % 10, H)),
                    "ds": pa.array(np.tile(grid, n_models)).cast(pa.timestamp("ns")),
                    "yhat": pa.array(rng.randint(0, 10**6, n_models * H).astype(np.int32))})
    frame = ps.ProphetScorer.convert_forecasts(Frame(tbl))
    cfg = lambda d, w: {"io": {"models": "unused", "forecasts": str(tmp_path / d)}, "forecast": {"periods": H, "frequency": "15min", "writer": w}}

    def rows(d):
        out = []
        for fn in sorted(os.listdir(tmp_path / d)):
            lines = open(tmp_path / d / fn, "rb").read().split(b"\n")
            assert lines[0] + b"\n" == ps._CSV_HEADER and lines[-1] == b""
            out += lines[1:-1]
        return out

    ps.ProphetScorer(cfg("arrow", "arrow")).write_forecasts(frame)
    ps.ProphetScorer(cfg("gpu1", "gpu")).write_forecasts(frame)
    assert os.listdir(tmp_path / "gpu1") == ["part-00000.csv"]
    assert open(tmp_path / "gpu1" / "part-00000.csv", "rb").read() == open(tmp_path / "arrow" / "part-00000.csv", "rb").read()
    monkeypatch.setattr(ps, "_GPU_ROWS_PER_PART", 5000)
    monkeypatch.setattr(ps, "_ROWS_PER_PART", 1000)
    ps.ProphetScorer(cfg("gpu3", "auto")).write_forecasts(frame)                 # above _ROWS_PER_PART: auto takes the GPU route
    assert sorted(os.listdir(tmp_path / "gpu3")) == ["part-00000-0000.csv", "part-00000-0001.csv", "part-00000-0002.csv"]
    assert rows("gpu3") == rows("arrow") and len(rows("arrow")) == n_models * H
    # frames the formatter does not cover keep the Arrow writer under auto and are refused under gpu
    tbl2 = tbl.append_column("yhat_lower", pa.array(np.zeros(n_models * H))).append_column("yhat_upper", pa.array(np.ones(n_models * H)))
    f2 = ps.ProphetScorer.convert_forecasts(Frame(tbl2))
    ps.ProphetScorer(cfg("iv", "auto")).write_forecasts(f2)
    assert len(os.listdir(tmp_path / "iv")) >= 1
    with pytest.raises(ValueError):
        ps.ProphetScorer(cfg("iv2", "gpu")).write_forecasts(f2)


def test_make_future_device_matches_host(gpu_ctx):
    """pb200_make_future_device = Prophet.make_future_dataframe(include_history=False) for a fixed-width frequency:
    last + (1..periods) * freq, the grid batched.make_future builds on the host (and pandas date_range in test_boundary)."""
    import torch
    from time_series_spark_b200 import batched
    last = np.array([0, 1_546_300_800 * 10**9, 1_615_851_900 * 10**9, -5 * 10**9], np.int64)
    for periods, freq in ((40, 15 * 60 * 10**9), (1, 7 * 86400 * 10**9), (672, 15 * 60 * 10**9)):
        got = batched.make_future_device(gpu_ctx, torch.from_numpy(last).cuda(), periods, freq).cpu().numpy()
        assert np.array_equal(got, batched.make_future(last, periods, freq))
