"""CPU tests of the oracle (oracle/prophet_oracle.py): analytic gradient vs finite differences,
the facts SURVEY.md section 8c derives from the reference's fixture, the committed golden
vectors, and an independent optimiser."""
import numpy as np
import pytest

from oracle import prophet_oracle as po


def _series(gi, dim):
    m = gi["dim_id"] == dim
    return gi["ds_ns"][m], gi["y"][m].astype(np.float64)


def test_fixture_facts(golden_input):
    """816 rows, dims 91 (410 rows, 4 duplicate rows) / 155 (406 rows); span 722 d 10:30
    => yearly off, weekly + daily on, K = 14, S = 25; cap = max(y) * 1.1 in double."""
    gi = golden_input
    assert gi["ds_ns"].size == 816 and set(np.unique(gi["dim_id"])) == {91, 155}
    opts = po.ProphetOptions()
    for dim, n, ymax, cap in ((91, 410, 94174, 103591.40000000001), (155, 406, 127322, 140054.2)):
        ds, y = _series(gi, dim)
        assert ds.size == n and y.max() == ymax
        p = po.prepare(ds, y, 0.0, y.max() * 1.1, opts)
        assert p.cap_value == cap
        assert [s.name for s in p.seasonalities] == ["weekly", "daily"]
        assert (p.K, p.S, p.y_scale) == (14, 25, float(ymax))
        assert p.t_scale_ns == (722 * 86400 + 10 * 3600 + 30 * 60) * 10**9
        assert np.allclose(p.cap, 1.1)
    ds, _ = _series(gi, 91)
    assert np.sum(np.diff(np.sort(ds)) == 0) == 4
    hist = int(np.floor(410 * 0.8))
    assert hist == 328
    idx = po.changepoint_indexes(410, opts)
    assert idx[0] == 13 and idx[1] == 26 and idx[-2] == 314 and idx[-1] == 327


@pytest.mark.parametrize("growth,mode", [("logistic", "multiplicative"), ("linear", "additive"),
                                         ("linear", "multiplicative"), ("logistic", "additive")])
def test_gradient_matches_finite_differences(golden_input, growth, mode):
    ds, y = _series(golden_input, 155)
    opts = po.ProphetOptions(growth=growth, seasonality_mode=mode)
    p = po.prepare(ds, y, 0.0, y.max() * 1.1, opts)
    rng = np.random.RandomState(3)
    th = po.initial_theta(p) + 0.1 * rng.randn(p.S + p.K + 3)
    err, f, g = po.neg_logp_grad(th, p)
    assert err == 0
    num = np.zeros_like(th)
    for i in range(th.size):
        h = 1e-6
        a, b = th.copy(), th.copy()
        a[i] += h
        b[i] -= h
        num[i] = (po.neg_logp_grad(a, p)[1] - po.neg_logp_grad(b, p)[1]) / (2 * h)
    assert np.max(np.abs(num - g) / (1 + np.abs(g))) < 1e-6


def test_oracle_reproduces_golden(golden_input, golden_oracle):
    """Regression pin of the oracle itself against tests/golden/fixture_751_oracle.npz."""
    go = golden_oracle
    opts = po.ProphetOptions()
    for dim in (91, 155):
        k = f"d{dim}_"
        ds, y = _series(golden_input, dim)
        p = po.prepare(ds, y, 0.0, float(go[k + "cap"]), opts)
        assert np.array_equal(p.t_change, go[k + "t_change"])
        assert np.allclose(po.initial_theta(p), go[k + "theta0"], rtol=0, atol=1e-15)
        for th, f, g in zip(go[k + "points"], go[k + "f"], go[k + "g"]):
            err, f2, g2 = po.neg_logp_grad(th, p)
            assert err == 0 and abs(f2 - f) <= 1e-9 * abs(f) and np.allclose(g2, g, rtol=1e-9, atol=1e-9)
        fr = po.fit(ds, y, 0.0, None, opts)
        assert fr.ret == int(go[k + "ret"]) == po.TERM_RELGRAD
        assert abs(fr.neg_logp - float(go[k + "neg_logp"])) < 1e-6
        fut = po.make_future_ns(fr.last_ds_ns, 40, 15 * 60 * 10**9)
        assert np.array_equal(fut, go[k + "future_ns"])
        pred = po.predict(fr, fut, 0.0, float(go[k + "cap32"]), opts)
        assert np.max(np.abs(pred["yhat"] - go[k + "yhat_future"])) < 1e-3 * float(go[k + "y_scale"])


def test_independent_optimiser_reaches_lower_or_equal_objective(golden_input):
    """scipy L-BFGS-B from the same start ends at an objective <= Stan's (Stan stops early on
    its relative-gradient rule); guards the objective/gradient against self-consistent bugs."""
    from scipy.optimize import minimize
    ds, y = _series(golden_input, 155)
    opts = po.ProphetOptions()
    p = po.prepare(ds, y, 0.0, y.max() * 1.1, opts)
    fr = po.fit(ds, y, 0.0, None, opts)
    res = minimize(lambda x: po.neg_logp_grad(x, p)[1:], po.initial_theta(p), jac=True, method="L-BFGS-B",
                   options=dict(maxiter=20000, maxfun=100000, ftol=1e-15, gtol=1e-10))
    assert res.fun <= fr.neg_logp + 1e-6
    assert fr.neg_logp - res.fun < 5.0        # same basin: Stan's early stop is within a few nats


def test_predict_pieces(golden_input):
    ds, y = _series(golden_input, 91)
    opts = po.ProphetOptions()
    fr = po.fit(ds, y, 0.0, None, opts)
    # in-sample prediction equals the Stan model's mean function at the optimum
    pr = po.predict(fr, fr.prep.ds_sorted, 0.0, fr.prep.cap_value, opts)
    th = fr.theta
    p = fr.prep
    kt = th[0] + p.A @ th[2:2 + p.S]
    assert np.all(np.isfinite(pr["yhat"])) and kt.shape == (p.T,)
    assert np.array_equal(po.scorer_epilogue(np.array([-3.7, 2.9, 0.2]), 0.0), np.array([0, 2, 0]))
    fut = po.make_future_ns(int(ds.max()), 3, 15 * 60 * 10**9)
    assert np.array_equal(np.diff(fut), [15 * 60 * 10**9] * 2) and fut[0] - ds.max() == 15 * 60 * 10**9


def test_constant_linear_shortcut_and_errors():
    ds = (np.arange(10) * 86400 * 10**9).astype(np.int64)
    fr = po.fit(ds, np.full(10, 5.0), opts=po.ProphetOptions(growth="linear"))
    assert fr.sigma_obs == 1e-9 and fr.iters == 0
    with pytest.raises(ValueError):
        po.fit(ds[:1], np.array([1.0]))
    with pytest.raises(ValueError):
        po.fit(ds, np.arange(10.0), floor=100.0, cap=50.0)
