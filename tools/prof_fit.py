"""Dev helper: one fit call on config-3 series, inputs resident in HBM (for ncu / event timing)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from time_series_spark_b200 import synth, batched, _lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = sys.argv[3] if len(sys.argv) > 3 else "c3"
ctx = L.Context(0)
if cfg == "c3":
    b = synth.config3(n=n); opts = batched.make_options()
elif cfg == "c2":
    b = synth.config2(n=n); opts = batched.make_options(growth="linear", yearly_seasonality=True)
else:
    b = synth.config4(n=n); opts = batched.make_options()
ds = torch.from_numpy(b.ds).cuda(); y = torch.from_numpy(b.y).cuda()
out = None
for r in range(reps):
    torch.cuda.synchronize(); t = time.time()
    out = batched.fit_batch_device(ctx, opts, ds, y, b.offsets, 0.0, 1.1, out=out)
    dt = time.time() - t
    print(f"{cfg} n={n} rep{r}: {dt*1e3:.2f} ms  {n/dt:.0f} series/s, evals/series {out.meta_i32[:,6].double().mean().item():.1f}")
