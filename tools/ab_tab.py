"""Dev helper: config-3 fit with and without the seasonal-table variants (same process, same inputs)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from time_series_spark_b200 import synth, batched, _lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
b = synth.config3(n=n); opts = batched.make_options()
ds = torch.from_numpy(b.ds).cuda(); y = torch.from_numpy(b.y).cuda()
res = {}
for name, env in (("tab", {}), ("notab", {"PB200_NO_TAB": "1"})):
    os.environ.pop("PB200_NO_TAB", None)
    os.environ.update(env)
    ctx = L.Context(0)
    out = None
    for r in range(reps):
        torch.cuda.synchronize(); t = time.time()
        out = batched.fit_batch_device(ctx, opts, ds, y, b.offsets, 0.0, 1.1, out=out)
        dt = time.time() - t
        print(f"{name} n={n} rep{r}: {dt*1e3:.2f} ms  {n/dt:.0f} series/s, evals/series {out.meta_i32[:,6].double().mean().item():.1f}", flush=True)
    print(name, "variant counts (rows: planes, rotation, week table, day table):", ctx.last_fit_variant_counts()[:, 6].tolist())
    res[name] = (out.meta_f64[:, 3].cpu().numpy().copy(), out.meta_i32[:, 4].cpu().numpy().copy())
    ctx.close()
fa, sa = res["tab"]; fb, sb = res["notab"]
rel = np.abs(fa - fb) / np.abs(fb)
print("objective at optimum, tab vs notab: median rel %.2e  p99 %.2e  max %.2e; identical %d/%d" %
      (np.median(rel), np.quantile(rel, 0.99), rel.max(), int((fa == fb).sum()), n))
print("status histogram tab:", dict(zip(*np.unique(sa, return_counts=True))), " notab:", dict(zip(*np.unique(sb, return_counts=True))))
