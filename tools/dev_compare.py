"""Dev helper (GPU box): GPU fit vs oracle on small samples of each config + a timing run."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from time_series_spark_b200 import synth, batched, _lib as L
from oracle import prophet_oracle as po

ctx = L.Context(0)

def compare(name, b, opts, oopts, nmax=16):
    b = b.take(0, min(nmax, b.n))
    t = time.time()
    fb = batched.fit_batch_host(ctx, opts, b.ds, b.y, b.offsets, 0.0, 1.1)
    dt = time.time() - t
    print(f"== {name}: n={b.n} gpu {dt*1e3:.1f} ms")
    for i in range(b.n):
        a, e = b.offsets[i], b.offsets[i + 1]
        fr = po.fit(b.ds[a:e], b.y[a:e].astype(float), opts=oopts)
        S = fr.prep.S; K = fr.prep.K
        pr = fb.params[i]
        st, it, ne = fb.meta_i32[i, 4], fb.meta_i32[i, 5], fb.meta_i32[i, 6]
        kd = abs(pr[0] - fr.k); md = abs(pr[1] - fr.m); sd = abs(pr[2] - fr.sigma_obs) / fr.sigma_obs
        dd = np.max(np.abs(pr[3:3 + S] - fr.delta)); bd = np.max(np.abs(pr[3 + fb.smax:3 + fb.smax + K] - fr.beta)) if fr.prep.seasonalities else 0
        print(f"  [{i}] st {st}/{fr.ret} it {it}/{fr.iters} ev {ne}/{fr.n_evals} f {fb.meta_f64[i,3]:.9f}/{fr.neg_logp:.9f} "
              f"dk {kd:.2e} dm {md:.2e} dsig {sd:.2e} ddelta {dd:.2e} dbeta {bd:.2e}")

o3 = batched.make_options()
compare("c3", synth.config3(n=16), o3, po.ProphetOptions(), 8)
o2 = batched.make_options(growth="linear", yearly_seasonality=True)
compare("c2", synth.config2(n=16), o2, po.ProphetOptions(growth="linear", yearly_seasonality=True), 8)
compare("c4", synth.config4(n=16), o3, po.ProphetOptions(), 8)

for n in (2048, 16384):
    b = synth.config3(n=n)
    for rep in range(2):
        t = time.time()
        fb = batched.fit_batch_host(ctx, o3, b.ds, b.y, b.offsets, 0.0, 1.1)
        dt = time.time() - t
        print(f"c3 n={n} rep{rep}: {dt:.3f}s  {n/dt:.0f} series/s  mean evals {fb.meta_i32[:,6].mean():.1f} mean iters {fb.meta_i32[:,5].mean():.1f}")
    st, cnt = np.unique(fb.meta_i32[:, 4], return_counts=True)
    print("   status hist", dict(zip(st.tolist(), cnt.tolist())))
b = synth.config4(n=65536)
for rep in range(2):
    t = time.time()
    fb = batched.fit_batch_host(ctx, o3, b.ds, b.y, b.offsets, 0.0, 1.1)
    dt = time.time() - t
    print(f"c4 n={b.n} rep{rep}: {dt:.3f}s  {b.n/dt:.0f} series/s  mean evals {fb.meta_i32[:,6].mean():.1f}")
st, cnt = np.unique(fb.meta_i32[:, 4], return_counts=True)
print("   status hist", dict(zip(st.tolist(), cnt.tolist())))
b = synth.config2()
for rep in range(2):
    t = time.time()
    fb = batched.fit_batch_host(ctx, o2, b.ds, b.y, b.offsets, 0.0, 1.1)
    dt = time.time() - t
    print(f"c2 n={b.n} rep{rep}: {dt:.3f}s  {b.n/dt:.0f} series/s  mean evals {fb.meta_i32[:,6].mean():.1f}")
st, cnt = np.unique(fb.meta_i32[:, 4], return_counts=True)
print("   status hist", dict(zip(st.tolist(), cnt.tolist())))
