"""profiles/fit_kernel_traffic.json from an ncu launch list of bench.py
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3
Usage: python tools/traffic_from_launches.py <launches.csv> <n_series per launch> [out.json]
Prints the share of GPU time per kernel and writes the dominant fit kernel's DRAM bytes per launch / per series,
labelled with the digest of the build it was captured on (bench.py compares it with the running build)."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, n_series = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "fit_kernel_traffic.json")
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"')) ]
hdr = rows[0]
ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.defaultdict(dict)
for r in rows[1:]:
    per[(r[iid], r[ik])][r[im]] = float(r[iv].replace(",", ""))
agg = collections.defaultdict(lambda: collections.Counter())
for (_, k), m in per.items():
    a = agg[k]
    a["n"] += 1
    a["ns"] += m.get("gpu__time_duration.sum", 0.0)
    a["rd"] += m.get("dram__bytes_read.sum", 0.0)
    a["wr"] += m.get("dram__bytes_write.sum", 0.0)
tot = sum(a["ns"] for a in agg.values())
print(f"{'kernel':90s} launches  share  ms/launch   rd MB/launch  wr MB/launch")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    print(f"{k[:90]:90s} {a['n']:7d} {100 * a['ns'] / tot:6.2f}% {a['ns'] / a['n'] / 1e6:9.3f} {a['rd'] / a['n'] / 1e6:12.1f} {a['wr'] / a['n'] / 1e6:12.1f}")
# the dominant fit kernel: launches that actually fitted the batch (not the empty-queue launches of other classes)
fits = {k: a for k, a in agg.items() if "fit_" in k}
top = max(fits, key=lambda k: fits[k]["ns"])
big = [m for (i, k), m in per.items() if k == top and m.get("gpu__time_duration.sum", 0.0) > 0.2 * fits[top]["ns"] / fits[top]["n"]]
rd = sum(m["dram__bytes_read.sum"] for m in big) / len(big)
wr = sum(m["dram__bytes_write.sum"] for m in big) / len(big)
sys.path.insert(0, ROOT)
try:
    from time_series_spark_b200.build import _sources_digest
    digest = _sources_digest()[:16]
except Exception:
    digest = None
d = {"kernel": top, "capture": f"{os.path.relpath(path, ROOT)}: mean of the {len(big)} launches of the kernel that fitted the batch, "
                               f"each over {n_series} series", "build_digest": digest, "n_series": n_series,
     "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_series": (rd + wr) / n_series,
     "algorithmic_bytes_per_series": 18056,
     "ms_per_launch_under_ncu": sum(m["gpu__time_duration.sum"] for m in big) / len(big) / 1e6,
     "note": "reads = ds / y once (17.3 KB per series, the algorithmic input); writes = the per-series workspace (y plane 11.5 KB + "
             "L-BFGS history 3.8 KB, rewritten for every series and written back from L2 once) + results"}
json.dump(d, open(out, "w"), indent=1)
print("wrote", out, json.dumps(d)[:300])
