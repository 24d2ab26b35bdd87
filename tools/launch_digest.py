"""Per-kernel digest of an ncu launch list (gpu_check.sh: --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv): launches, average duration, average DRAM bytes, share of the summed time.  usage: launch_digest.py launches.csv [traffic.json]"""
import csv
import json
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
col = {k: i for i, k in enumerate(hdr)}
per = defaultdict(lambda: defaultdict(float))
units = {}
for r in rows[1:]:
    per[r[col["ID"]]]["name"] = r[col["Kernel Name"]]
    v = float(r[col["Metric Value"]].replace(",", ""))
    u = r[col["Metric Unit"]]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    per[r[col["ID"]]][r[col["Metric Name"]]] = v * scale


def short(name):
    m = re.match(r"(?:void )?(?:chiml::)?(k_\w+)(?:<(\d))?", name)
    if not m:
        return name
    k = m.group(1)
    if k in ("k_fast", "k_uniform", "k_general", "k_uniform_rows"):
        return f"{k}<{'E' if m.group(2) == '1' else 'H'}>"
    return k            # argument lists dropped


agg = defaultdict(lambda: [0, 0.0, 0.0])
for d in per.values():
    a = agg[short(d["name"])]
    a[0] += 1
    a[1] += d["gpu__time_duration.sum"]
    a[2] += d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':22s} {'launches':>8s} {'avg ms':>8s} {'share':>6s} {'DRAM GB/launch':>15s} {'DRAM GB/s':>10s}")
traffic = {}
for k, (n, ms, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:22s} {n:8d} {ms / n:8.3f} {ms / tot:6.1%} {b / n / 1e9:15.3f} {b / ms / 1e6:10.0f}")
    traffic[k] = round(b / n, -6)
if len(sys.argv) > 2:
    traffic = {"_comment": f"dram__bytes_read.sum + dram__bytes_write.sum per launch, bytes, from profiles/{sys.argv[1].split('/')[-1]} "
                           "(ncu launch list of the default bench workload, 2048x256x1024, one B200)", **traffic}
    json.dump(traffic, open(sys.argv[2], "w"), indent=1)
