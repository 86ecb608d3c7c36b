"""Summarise an `ncu --page source --csv` dump: stall reasons and the hottest SASS lines.

usage: ncu -i X.ncu-rep --page source --csv | python profiles/ncu_source_summary.py [kernel_index] [top_n]
"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = heads[which]
end = heads[which + 1] - 1 if which + 1 < len(heads) else len(rows)
hdr = rows[hi]
body = [r for r in rows[hi + 1:end] if len(r) == len(hdr) and r[0].startswith("0x")]
ix = {h: i for i, h in enumerate(hdr)}
print(rows[hi - 1][:2])
tot = sum(int(r[ix["# Samples"]]) for r in body)
ex = sum(int(r[ix["Instructions Executed"]]) for r in body)
print(f"samples={tot} sass_lines={len(body)} warp_instructions_executed={ex}")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in body) for h in stall_cols}
for h, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
    print(f"  {h:24s} {v:8d} {100 * v / max(tot, 1):5.1f}%")
print("--- hottest SASS lines (samples, executed, instruction)")
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
    print(f"  {r[ix['# Samples']]:>6s} {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:100]}")
