"""Summarises the SASS page of an ncu report (`ncu -i X.ncu-rep --page source --csv`): stall reasons overall and the
instructions that collect the most samples, with a few lines of context.   stall_summary.py src.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: j for j, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print(f"total samples {tot}  instructions {len(body)}")
print(sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:10])
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top_n]
for i in order:
    r = body[i]
    n = int(r[ix["# Samples"]] or 0)
    why = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{r[0][-5:]} {n:6d} {100.0 * n / tot:5.1f}%  {r[ix['Source']][:70]:70s} {why}  exec={r[ix['Instructions Executed']]}")
