"""Dev helper: forecast-frame writer (convert_forecasts + write_forecasts) on a config-#5-shaped frame, Arrow vs GPU row formatter.
    python tools/bench_writer.py <n_models> [arrow|gpu|auto]      # n_models x 672 rows"""
import os, shutil, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pyarrow as pa, torch  # noqa: F401  (torch first: its import is not what is being timed)
from time_series_spark_b200.frame import Frame
from time_series_spark_b200.jobs.prophet_scorer import ProphetScorer

n_models, H = int(sys.argv[1]), 672
writer = sys.argv[2] if len(sys.argv) > 2 else "auto"
n = n_models * H
t = pa.table({"series_id": pa.array(np.repeat(np.arange(n_models, dtype=np.int32) // 100, H)),
              "dim_id": pa.array(np.repeat(np.arange(n_models, dtype=np.int32) % 100, H)),
              "ds": pa.array(np.tile(1_650_000_000_000_000_000 + 900 * 10**9 * np.arange(H, dtype=np.int64), n_models)).cast(pa.timestamp("ns")),
              "yhat": pa.array(np.random.RandomState(0).randint(0, 100000, n).astype(np.int32))})
out = "/tmp/fc_bench_" + writer
sc = ProphetScorer({"io": {"forecasts": out}, "forecast": {"writer": writer}})
for rep in range(2):
    shutil.rmtree(out, ignore_errors=True)
    t0 = time.time(); f = ProphetScorer.convert_forecasts(Frame(t)); t1 = time.time(); sc.write_forecasts(f); t2 = time.time()
    sz = sum(os.path.getsize(os.path.join(out, x)) for x in os.listdir(out))
    print(f"{writer}: {n / 1e6:.1f} M rows rep{rep}: convert {t1 - t0:.2f} s, write {t2 - t1:.2f} s, {sz / 1e6:.0f} MB in {len(os.listdir(out))} files", flush=True)
