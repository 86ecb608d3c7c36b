"""Aggregate an ncu launch list with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum by kernel.
usage: python profiles/ncu_epoch_summary.py launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
idi = hdr.index("ID")
per = collections.defaultdict(dict)
names = {}
for r in rows[hdr_i + 1:]:
    if len(r) != len(hdr):
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "byte": 1e-9, "Kbyte": 1e-6,
             "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1.0)
    per[r[idi]][r[mi]] = v * scale
    names[r[idi]] = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("d2d::", "")
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
for i, m in per.items():
    a = agg[names[i]]
    a[0] += m.get("gpu__time_duration.sum", 0.0)
    a[1] += m.get("dram__bytes_read.sum", 0.0)
    a[2] += m.get("dram__bytes_write.sum", 0.0)
    a[3] += 1
total = sum(a[0] for a in agg.values())
print(f"total kernel time {total:.2f} ms")
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:16]:
    gbs = (a[1] + a[2]) / (a[0] * 1e-3) if a[0] else 0.0
    print(f"{100 * a[0] / total:6.2f}% {a[0]:9.2f} ms n={a[3]:4d} avg {1e3 * a[0] / a[3]:9.1f} us  dram rd {a[1]:7.2f} GB wr {a[2]:7.2f} GB "
          f"({gbs:5.0f} GB/s)  {k[:60]}")
