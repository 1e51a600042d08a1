"""Dev helper: time predict (deterministic and with 1000-draw intervals) on fitted config-3 models."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from time_series_spark_b200 import synth, batched, _lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = L.Context(0)
b = synth.config3(n=n); opts = batched.make_options()
ds = torch.from_numpy(b.ds).cuda(); y = torch.from_numpy(b.y).cuda()
fb = batched.fit_batch_device(ctx, opts, ds, y, b.offsets, 0.0, 1.1)
H = 672
last = torch.from_numpy(b.ds[b.offsets[1:] - 1].copy()).cuda()
fut = (last[:, None] + (15 * 60 * 10**9) * torch.arange(1, H + 1, device="cuda", dtype=torch.int64)[None, :]).contiguous()
fl = torch.zeros(n, dtype=torch.float64, device="cuda"); cap = fb.meta_f64[:, 2].float().double().contiguous()
for mc in (False, True):
    o = batched.make_options(uncertainty_samples=1000 if mc else 0)
    for r in range(2):
        torch.cuda.synchronize(); t = time.time()
        fc = batched.predict_batch_device(ctx, o, fb, fut, fl, cap, seed=3, intervals=mc)
        dt = time.time() - t
    print(f"predict mc={mc}: {dt*1e3:.2f} ms, {n*H/dt/1e6:.1f} M points/s")
    if mc:
        print("  lower<yhat<upper:", bool(torch.all(fc.yhat_lower < fc.yhat).item() and torch.all(fc.yhat < fc.yhat_upper).item()),
              " mean width / (sigma*y_scale):", float(((fc.yhat_upper - fc.yhat_lower).mean(1) / (fb.params[:, 2] * fb.meta_f64[:, 0])).mean()))
