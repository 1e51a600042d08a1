"""Dev helper: pb200_fit_host (pinned host buffers, copies inside the call) on config-3 series under several
pb200_create environments, e.g.   python tools/ab_host.py 50000 3 c1:PB200_HOST_CHUNKS=1 c2:PB200_HOST_CHUNKS=2 c4:"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from time_series_spark_b200 import synth, batched, _lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
specs = sys.argv[3:] or ["base:"]
b = synth.config3(n=n); opts = batched.make_options()
ds = torch.from_numpy(b.ds).pin_memory().numpy(); y = torch.from_numpy(b.y).pin_memory().numpy()
for spec in specs:
    name, _, kv = spec.partition(":")
    for k in ("PB200_HOST_CHUNKS", "PB200_GROUP", "PB200_GROUP_MIN"):
        os.environ.pop(k, None)
    for item in filter(None, kv.split(",")):
        k, _, v = item.partition("=")
        os.environ[k] = v
    ctx = L.Context(0)
    for r in range(reps + 1):
        torch.cuda.synchronize(); t = time.perf_counter()
        out = batched.fit_batch_host(ctx, opts, ds, y, b.offsets, 0.0, 1.1)
        dt = time.perf_counter() - t
        if r:
            print(f"{name} n={n} rep{r}: {dt*1e3:.2f} ms  {n/dt:.0f} series/s", flush=True)
    ctx.close()
