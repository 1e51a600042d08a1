"""Dev helper: how much does the ORDER in which series are popped matter?  Fits config #3 once, then refits the same series
permuted by their true evaluation counts (descending = longest-processing-time-first with perfect knowledge, ascending = the
worst case, random) with the cost key switched off (PB200_QKEY_CV=0: equal-length series share one bin, so the queue is in
index order).  Upper bound of what a better cost predictor could buy.
    PB200_QKEY_CV=0 python tools/ab_order.py 50000"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from time_series_spark_b200 import synth, batched, _lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
b = synth.config3(n=n); opts = batched.make_options()
T = int(b.offsets[1] - b.offsets[0])
ctx = L.Context(0)
ds = torch.from_numpy(b.ds).cuda(); y = torch.from_numpy(b.y).cuda()
out = batched.fit_batch_device(ctx, opts, ds, y, b.offsets, 0.0, 1.1)
ev = out.meta_i32[:, 6].cpu().numpy()
print(f"evals: mean {ev.mean():.1f} max {ev.max()} p99 {np.percentile(ev, 99):.0f}; cost key weight {os.environ.get('PB200_QKEY_CV', '2 (default)')}")
rng = np.random.RandomState(0)
orders = {"as generated": np.arange(n), "random": rng.permutation(n), "longest first (true evals)": np.argsort(-ev, kind="stable"),
          "shortest first": np.argsort(ev, kind="stable")}
for name, o in orders.items():
    dsp = torch.from_numpy(b.ds.reshape(n, T)[o].reshape(-1).copy()).cuda()
    yp = torch.from_numpy(b.y.reshape(n, T)[o].reshape(-1).copy()).cuda()
    res = None
    ts = []
    for r in range(3):
        torch.cuda.synchronize(); t = time.time()
        res = batched.fit_batch_device(ctx, opts, dsp, yp, b.offsets, 0.0, 1.1, out=res)
        ts.append(time.time() - t)
    same = bool((res.meta_i32[:, 6].cpu().numpy() == ev[o]).all())
    print(f"{name:28s}: {min(ts[1:])*1e3:.1f} ms   (same evaluation counts as the first fit: {same})", flush=True)
