"""Dynamic opcode mix of one kernel from `ncu --page source --csv` (warp-level executed counts).

usage: ncu -i X.ncu-rep --page source --csv | python profiles/ncu_opcode_mix.py [kernel_index] [units_per_launch]
"""
import collections
import csv
import sys

rows = list(csv.reader(sys.stdin))
which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = heads[which]
end = heads[which + 1] - 1 if which + 1 < len(heads) else len(rows)
hdr = rows[hi]
body = [r for r in rows[hi + 1:end] if len(r) == len(hdr) and r[0].startswith("0x")]
ix = {h: i for i, h in enumerate(hdr)}
c = collections.Counter()
for r in body:
    toks = r[ix["Source"]].strip().split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    c[op.split(".")[0]] += int(r[ix["Instructions Executed"]])
tot = sum(c.values())
print(rows[hi - 1][1], "kernels in report:", len(heads))
for k, v in c.most_common(22):
    extra = f"  {v * 32 / units:7.1f} thread-instr/unit" if units else ""
    print(f"{k:10s} {v:11d} {100 * v / tot:5.1f}%{extra}")
print("total", tot, f"{tot * 32 / units:.1f} thread-instr/unit" if units else "")
