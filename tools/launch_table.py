"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total and mean device time,
share.  python tools/launch_table.py gpurun_out/x.csv [skip_first_n_launches]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
tot = 0.0
for r in rows[hdr + 1 + skip:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    n = r[ki].replace("<unnamed>::", "").replace("void ", "")[:70]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print("%-72s %8s %12s %10s %7s" % ("kernel", "launches", "total_us", "mean_us", "share"))
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-72s %8d %12.1f %10.2f %6.1f%%" % (n, c, t, t / c, 100 * t / tot))
print("total %.1f us over %d launches" % (tot, sum(c for c, _ in agg.values())))
