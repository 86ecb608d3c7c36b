"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/ncu_launch_summary.py launches.csv [skip_first_n]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
    tot[name] += float(r[-1])
    cnt[name] += 1
total = sum(tot.values())
print(f"{len(rows)} launches, {total / 1e6:.3f} ms of kernel time")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:25]:
    print(f"{100 * v / total:6.2f}%  {v / 1e6:9.3f} ms  n={cnt[k]:5d}  avg {v / cnt[k] / 1e3:9.1f} us  {k[:110]}")
