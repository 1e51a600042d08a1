"""Synthetic workloads of BASELINE.json `configs`, as ragged batches
in the reference's input schema (series_id, dim_id, timestamp, quantity)
(MODEL_INPUT_SCHEMA, reference src/jobs/prophet_modeler.py:12-17).

Generation is block-wise (1024 series per block, seeded by (seed, block)) so that any
rank can produce exactly its own shard without touching the others.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

NS_MIN = 60 * 10**9
NS_DAY = 86400 * 10**9
BLOCK = 1024


@dataclass
class RaggedBatch:
    series_id: np.ndarray   # [N] int32
    dim_id: np.ndarray      # [N] int32
    offsets: np.ndarray     # [N+1] int64
    ds: np.ndarray          # [R] int64 ns since epoch, ascending within a series
    y: np.ndarray           # [R] int32 (quantity)

    @property
    def n(self) -> int:
        return self.offsets.size - 1

    def take(self, lo: int, hi: int) -> "RaggedBatch":
        a, b = int(self.offsets[lo]), int(self.offsets[hi])
        return RaggedBatch(self.series_id[lo:hi], self.dim_id[lo:hi], self.offsets[lo:hi + 1] - a,
                           self.ds[a:b], self.y[a:b])


def _epoch_ns(s: str) -> int:
    return int(np.datetime64(s, "ns").astype(np.int64))


def _blocks(lo: int, hi: int):
    b = lo // BLOCK
    while b * BLOCK < hi:
        s, e = max(lo, b * BLOCK), min(hi, (b + 1) * BLOCK)
        yield b, s - b * BLOCK, e - b * BLOCK
        b += 1


def config3(n: int = 50_000, T: int = 1440, seed: int = 2024, lo: int = 0, hi: int | None = None) -> RaggedBatch:
    """Config #3 (headline): N x 1440 15-min points from 2021-03-01, saturating level x daily
    (2 harmonics) x weekly (1 harmonic) profile, 5 % multiplicative noise, int32 >= 1.
    Keys: series_id = i // 100, dim_id = i % 100.  Fit with the reference defaults
    (logistic growth, multiplicative seasonality, floor 0, cap_multiplier 1.1)."""
    hi = n if hi is None else hi
    start = _epoch_ns("2021-03-01T00:00:00")
    step = 15 * NS_MIN
    grid = start + step * np.arange(T, dtype=np.int64)
    days = (grid - start) / NS_DAY
    u = np.linspace(0.0, 1.0, T)
    ys = []
    for b, s, e in _blocks(lo, hi):
        rng = np.random.default_rng([seed, b])
        L = np.exp(rng.uniform(np.log(1e3), np.log(1e5), BLOCK))
        r = rng.uniform(2.0, 10.0, BLOCK) * rng.choice([-1.0, 1.0], BLOCK, p=[0.3, 0.7])
        t0 = rng.uniform(0.2, 0.8, BLOCK)
        a1, a2 = rng.uniform(0.05, 0.4, BLOCK), rng.uniform(0.0, 0.15, BLOCK)
        p1, p2, p3 = (rng.uniform(0, 2 * np.pi, BLOCK) for _ in range(3))
        w1 = rng.uniform(0.0, 0.2, BLOCK)
        noise = rng.normal(0.0, 0.05, (BLOCK, T))
        sl = slice(s, e)
        level = L[sl, None] * (0.25 + 0.75 / (1.0 + np.exp(-r[sl, None] * (u[None, :] - t0[sl, None]))))
        daily = 1.0 + a1[sl, None] * np.sin(2 * np.pi * days[None, :] + p1[sl, None]) \
            + a2[sl, None] * np.sin(4 * np.pi * days[None, :] + p2[sl, None])
        weekly = 1.0 + w1[sl, None] * np.sin(2 * np.pi * days[None, :] / 7.0 + p3[sl, None])
        v = level * daily * weekly * (1.0 + noise[sl])
        ys.append(np.maximum(np.rint(v), 1.0).astype(np.int32))
    y = np.concatenate(ys, axis=0) if ys else np.zeros((0, T), np.int32)
    m = hi - lo
    idx = np.arange(lo, hi)
    return RaggedBatch((idx // 100).astype(np.int32), (idx % 100).astype(np.int32),
                       (np.arange(m + 1, dtype=np.int64) * T), np.tile(grid, m), y.reshape(-1))


def config2(n: int = 1000, T: int = 365, seed: int = 1234, lo: int = 0, hi: int | None = None) -> RaggedBatch:
    """Config #2: N x 365 daily points from 2018-01-01, linear trend with one slope break,
    yearly + weekly seasonality, additive noise 3 % of base.  Fit with growth='linear',
    yearly_seasonality=True (auto would disable it: 364 d < 730 d)."""
    hi = n if hi is None else hi
    start = _epoch_ns("2018-01-01T00:00:00")
    grid = start + NS_DAY * np.arange(T, dtype=np.int64)
    d = np.arange(T, dtype=np.float64)
    u = d / (T - 1)
    ys = []
    for b, s, e in _blocks(lo, hi):
        rng = np.random.default_rng([seed, b])
        base = np.exp(rng.uniform(np.log(1e2), np.log(1e5), BLOCK))
        slope = rng.uniform(-0.3, 0.8, BLOCK)
        brk = rng.uniform(0.2, 0.8, BLOCK)
        dslope = rng.uniform(-0.5, 0.5, BLOCK)
        ph1, ph2 = rng.uniform(0, 2 * np.pi, BLOCK), rng.uniform(0, 2 * np.pi, BLOCK)
        noise = rng.normal(0.0, 0.03, (BLOCK, T))
        sl = slice(s, e)
        trend = 1.0 + slope[sl, None] * u[None, :] + dslope[sl, None] * np.maximum(u[None, :] - brk[sl, None], 0.0)
        seas = 1.0 + 0.1 * np.sin(2 * np.pi * d[None, :] / 365.25 + ph1[sl, None]) \
            + 0.05 * np.sin(2 * np.pi * d[None, :] / 7.0 + ph2[sl, None])
        v = base[sl, None] * (trend * seas + noise[sl])
        ys.append(np.maximum(np.rint(v), 1.0).astype(np.int32))
    y = np.concatenate(ys, axis=0) if ys else np.zeros((0, T), np.int32)
    m = hi - lo
    idx = np.arange(lo, hi)
    return RaggedBatch((idx // 100).astype(np.int32), (idx % 100).astype(np.int32),
                       (np.arange(m + 1, dtype=np.int64) * T), np.tile(grid, m), y.reshape(-1))


# Config-#4 series (default seed) on which Stan's L-BFGS ends in a line-search failure under the default tolerances
# (found with the C oracle over the whole 500k batch -- and none in 1.6 M series of four other seeds): the case
# fbprophet 0.5 answers with its Newton retry.  Used by tests/test_gpu_optimiser.py.
CONFIG4_LSFAIL_IDS = (148912,)


def config4(n: int = 500_000, seed: int = 4321, lo: int = 0, hi: int | None = None,
            tmin: int = 48, tmax: int = 96) -> RaggedBatch:
    """Config #4: N short ragged series, T_i ~ U{48..96}, 15-min spacing (span < 2 days so
    every auto seasonality is off, K = 1 dummy column), >= 50 dim_id per series_id."""
    hi = n if hi is None else hi
    start = _epoch_ns("2022-06-01T00:00:00")
    step = 15 * NS_MIN
    ds_l, y_l, len_l = [], [], []
    u = np.arange(tmax, dtype=np.float64)
    for b, s, e in _blocks(lo, hi):
        rng = np.random.default_rng([seed, b])
        Ts = rng.integers(tmin, tmax + 1, BLOCK)
        base = np.exp(rng.uniform(np.log(50.0), np.log(5e4), BLOCK))
        slope = rng.uniform(-0.5, 1.0, BLOCK)
        curve = rng.uniform(-0.5, 0.5, BLOCK)
        shift = rng.integers(0, 96, BLOCK)
        noise = rng.normal(0.0, 0.04, (BLOCK, tmax))
        for i in range(s, e):
            T = int(Ts[i])
            x = u[:T] / (T - 1)
            v = base[i] * (1.0 + slope[i] * x + curve[i] * x * x + noise[i, :T])
            y_l.append(np.maximum(np.rint(v), 1.0).astype(np.int32))
            ds_l.append(start + step * (int(shift[i]) + np.arange(T, dtype=np.int64)))
            len_l.append(T)
    m = hi - lo
    idx = np.arange(lo, hi)
    offs = np.zeros(m + 1, np.int64)
    np.cumsum(np.asarray(len_l, dtype=np.int64), out=offs[1:])
    return RaggedBatch((idx // 64).astype(np.int32), (idx % 64).astype(np.int32), offs,
                       np.concatenate(ds_l) if ds_l else np.zeros(0, np.int64),
                       np.concatenate(y_l) if y_l else np.zeros(0, np.int32))
