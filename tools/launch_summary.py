"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for the LAST step."""
import collections
import csv
import re
import sys

path = sys.argv[1]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 140
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
# skip torch's own kernels (copies/fills) when counting the step boundary: use only our kernels
ours = [r for r in rows if "unnamed" in r["Kernel Name"]]
last = ours[-per_step:]
agg = collections.OrderedDict()
tot = 0.0
for r in last:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")[:60]
    v = float(r["Metric Value"].replace(",", "")) / 1e6
    grid = r["Grid Size"]
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += v
    a[2] = max(a[2], v)
    tot += v
print(f"# last {len(last)} launches of our kernels: total {tot:.3f} ms (cold-cache serialised ncu timing)")
print(f"# {'ms':>9} {'share':>6} {'n':>4} {'max ms':>8}  kernel")
for k, (n, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.3f} {100*t/tot:5.1f}% {n:4d} {mx:8.3f}  {k}")
