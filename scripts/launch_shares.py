"""Per-kernel launch counts, mean duration and share of the total from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`).  usage: launch_shares.py X.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi and r[vi].replace(",", "").replace(".", "").isdigit():
        agg.setdefault(r[ki].split("(")[0].replace("void ", ""), []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print("kernel,launches,mean_us,share")
for k, v in agg.items():
    print("%s,%d,%.1f,%.4f" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / tot))
