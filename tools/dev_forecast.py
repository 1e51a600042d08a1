"""Dev helper (GPU box): forecast-level discrepancy GPU vs oracle, and oracle self-sensitivity."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from time_series_spark_b200 import synth, batched, _lib as L
from oracle import prophet_oracle as po

ctx = L.Context(0)
H = 96
freq = 15 * 60 * 10**9

def run(name, b, opts, oopts, n, freq):
    b = b.take(0, n)
    fb = batched.fit_batch_host(ctx, opts, b.ds, b.y, b.offsets, 0.0, 1.1)
    last = np.array([b.ds[b.offsets[i + 1] - 1] for i in range(n)])
    fut = batched.make_future(last, H, freq)
    cap32 = fb.meta_f64[:, 2].astype(np.float32).astype(np.float64)
    fc = batched.predict_batch_host(ctx, opts, fb, fut, np.zeros(n), cap32, seed=1, intervals=True)
    rel_fut, rel_self, rel_pred_same = [], [], []
    for i in range(n):
        a, e = b.offsets[i], b.offsets[i + 1]
        ds, y = b.ds[a:e], b.y[a:e].astype(float)
        fr = po.fit(ds, y, opts=oopts)
        pr = po.predict(fr, fut[i], 0.0, cap32[i], oopts)
        rel_fut.append(np.max(np.abs(pr["yhat"] - fc.yhat[i])) / fr.prep.y_scale)
        # oracle predict with the GPU's parameters: isolates the predict kernel
        S, K = fr.prep.S, fr.prep.K
        fr2 = po.FitResult(prep=fr.prep, k=fb.params[i, 0], m=fb.params[i, 1], delta=fb.params[i, 3:3 + S].copy(),
                           sigma_obs=fb.params[i, 2], beta=fb.params[i, 3 + fb.smax:3 + fb.smax + K].copy(),
                           theta=None, neg_logp=0, iters=0, n_evals=0, ret=0)
        pr2 = po.predict(fr2, fut[i], 0.0, cap32[i], oopts)
        rel_pred_same.append(np.max(np.abs(pr2["yhat"] - fc.yhat[i])) / fr.prep.y_scale)
        # oracle self-sensitivity: y perturbed in the last bit of the scaled value (1 ulp of one point)
        y2 = y.copy(); y2[len(y2) // 2] = np.nextafter(y2[len(y2) // 2], np.inf)
        fr3 = po.fit(ds, y2, opts=oopts)
        pr3 = po.predict(fr3, fut[i], 0.0, cap32[i], oopts)
        rel_self.append(np.max(np.abs(pr3["yhat"] - pr["yhat"])) / fr.prep.y_scale)
        if i < 3:
            rng = np.random.RandomState(5)
            un = po.predict_uncertainty(fr2, fut[i], pr2, rng, oopts)
            w_o = un["yhat_upper"] - un["yhat_lower"]; w_g = fc.yhat_upper[i] - fc.yhat_lower[i]
            print(f"   mc[{i}] width oracle {w_o.mean():.2f} gpu {w_g.mean():.2f}; lower diff/sd {np.mean(fc.yhat_lower[i]-un['yhat_lower'])/(fr2.sigma_obs*fr.prep.y_scale):.3f} upper {np.mean(fc.yhat_upper[i]-un['yhat_upper'])/(fr2.sigma_obs*fr.prep.y_scale):.3f}")
    f = lambda v: f"med {np.median(v):.2e} p90 {np.percentile(v,90):.2e} max {np.max(v):.2e}"
    print(f"== {name} n={n}: forecast max|dyhat|/y_scale GPU-vs-oracle: {f(rel_fut)}")
    print(f"     predict kernel alone (same params):            {f(rel_pred_same)}")
    print(f"     oracle vs oracle(1-ulp-perturbed input):        {f(rel_self)}")

run("c3", synth.config3(n=64), batched.make_options(), po.ProphetOptions(), 48, freq)
run("c2", synth.config2(n=64), batched.make_options(growth="linear", yearly_seasonality=True),
    po.ProphetOptions(growth="linear", yearly_seasonality=True), 48, 86400 * 10**9)
run("c4", synth.config4(n=64), batched.make_options(), po.ProphetOptions(), 48, freq)
