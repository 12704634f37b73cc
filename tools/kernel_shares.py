"""Per-kernel share of GPU time from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: kernel_shares.py launches.csv "<comment line>" > shares.csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iu = hdr.index("Metric Unit")
tot = OrderedDict()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("opb::(anonymous namespace)::", "").replace("void ", "")
    us = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[iu], 1.0)
    a = tot.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
total = sum(v[1] for v in tot.values())
print("# " + (sys.argv[2] if len(sys.argv) > 2 else ""))
print("kernel,launches,total_us,share_pct")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.1f,%.1f" % (k, n, us, 100 * us / total))
