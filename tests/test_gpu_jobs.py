"""GPU end-to-end tests of the drop-in jobs: the reference's own test flow
(tests/unit/prophet_modeler_test.py:59-75, tests/unit/prophet_scorer_test.py:83-114) with Spark
replaced by the batched operators, on the reference's fixture (config #1)."""
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


def test_make_future_device_matches_host(gpu_ctx):
    """pb200_make_future_device = Prophet.make_future_dataframe(include_history=False) for a fixed-width frequency:
    last + (1..periods) * freq, the grid batched.make_future builds on the host (and pandas date_range in test_boundary)."""
    import torch
    from time_series_spark_b200 import batched
    last = np.array([0, 1_546_300_800 * 10**9, 1_615_851_900 * 10**9, -5 * 10**9], np.int64)
    for periods, freq in ((40, 15 * 60 * 10**9), (1, 7 * 86400 * 10**9), (672, 15 * 60 * 10**9)):
        got = batched.make_future_device(gpu_ctx, torch.from_numpy(last).cuda(), periods, freq).cpu().numpy()
        assert np.array_equal(got, batched.make_future(last, periods, freq))
