"""Print per-launch duration / tensor-pipe activity from an `ncu --csv` log (one line per kernel launch)."""
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
h = rows[0]
ni, vi, ii = h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
d = {}
for r in rows[1:]:
    d.setdefault(r[ii], {})[r[ni]] = r[vi]
tot = 0.0
for k, v in d.items():
    t = float(v.get("gpu__time_duration.sum", "0").replace(",", ""))
    tot += t
    print("%3s  %10.1f us   tensor-active %s %%" % (k, t / 1e3, v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "-")))
print("total %.3f ms" % (tot / 1e6))
