"""Dev / acceptance helper: end-to-end modeler + scorer jobs on a synthetic hive-partitioned CSV input,
optionally under torchrun (one rank per GPU).  Usage:
    python tools/e2e_jobs.py make <dir> <n_series>        # write input + configs
    torchrun ... -m time_series_spark_b200.modeler_driver <dir>/modeler.yaml ; same for scorer
    python tools/e2e_jobs.py check <dir> <n_series> <periods>
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import yaml


def make(d, n):
    from time_series_spark_b200 import synth
    b = synth.config3(n=n, T=480)
    os.makedirs(d, exist_ok=True)
    inp = os.path.join(d, "input")
    ts = b.ds.astype("datetime64[ns]").astype("datetime64[s]").astype(str)
    for sid in np.unique(b.series_id):
        p = os.path.join(inp, f"series_id={sid}")
        os.makedirs(p, exist_ok=True)
        with open(os.path.join(p, "part-0.csv"), "w") as f:
            for i in np.flatnonzero(b.series_id == sid):
                a, e = b.offsets[i], b.offsets[i + 1]
                dim = b.dim_id[i]
                f.write("".join(f"{dim},{t.replace('T', ' ')},{q}\n" for t, q in zip(ts[a:e], b.y[a:e])))
    yaml.safe_dump({"io": {"input": inp, "models": os.path.join(d, "models")},
                    "model": {"floor": 0, "cap_multiplier": 1.1}}, open(os.path.join(d, "modeler.yaml"), "w"))
    yaml.safe_dump({"io": {"models": os.path.join(d, "models"), "forecasts": os.path.join(d, "forecasts")},
                    "forecast": {"periods": 96, "frequency": "15min"}}, open(os.path.join(d, "scorer.yaml"), "w"))


def check(d, n, periods):
    import pyarrow.dataset as pads
    m = pads.dataset(os.path.join(d, "models"), format="parquet").to_table()
    f = pads.dataset(os.path.join(d, "forecasts"), format="csv").to_table()
    parts_m = sorted(os.listdir(os.path.join(d, "models")))
    parts_f = sorted(os.listdir(os.path.join(d, "forecasts")))
    print("model rows", m.num_rows, "parts", parts_m, "| forecast rows", f.num_rows, "parts", parts_f)
    assert m.num_rows == n and f.num_rows == n * periods
    keys = set(zip(m["series_id"].to_pylist(), m["dim_id"].to_pylist()))
    assert len(keys) == n
    assert f.column_names == ["created_timestamp", "series_id", "dim_id", "forecast_date", "forecast_timestamp", "forecast_quantity"]
    print("e2e jobs OK")


if __name__ == "__main__":
    if sys.argv[1] == "make":
        make(sys.argv[2], int(sys.argv[3]))
    else:
        check(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
