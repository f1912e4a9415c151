"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

for f in sys.argv[1:]:
    with open(f) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v *= {"us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}.get(u, 1.0)
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    print(f, "total %.3f ms" % (T / 1e6))
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print("  %-28s n=%5d  %9.3f ms  %5.1f%%  avg %9.1f us" % (k, cnt[k], v / 1e6, 100 * v / T, v / cnt[k] / 1e3))
