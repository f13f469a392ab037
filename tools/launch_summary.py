"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel for the last P-frame."""
import collections
import csv
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ix = {h: j for j, h in enumerate(hdr)}
seq = []
for r in rows[start:]:
    if len(r) < len(hdr):
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("dmc::", "")
    seq.append((name, float(r[ix["Metric Value"]].replace(",", "")) / 1e3))
# a P frame starts at the last nchw_to_s3 launch (head of prog_head_p) and ends at k_finalize_bpp
ends = [i for i, (n, _) in enumerate(seq) if n.startswith("k_finalize_bpp")]
starts = [i for i, (n, _) in enumerate(seq) if n.startswith("k_nchw_to_s3")]
if len(ends) >= 1 and starts:
    e = ends[-1]
    s = max(i for i in starts if i < e)
    frame = seq[s:e + 1]
else:
    frame = seq
tot = sum(v for _, v in frame)
d = collections.defaultdict(lambda: [0, 0.0])
for n, v in frame:
    d[n][0] += 1
    d[n][1] += v
print(f"{len(frame)} launches in the frame, {tot:.1f} us summed (ncu per-launch times: serialised, cold)")
for k, (c, v) in sorted(d.items(), key=lambda kv: -kv[1][1]):
    print(f"{v:9.1f} us {100 * v / tot:5.1f}%  n={c:3d}  avg={v / c:7.1f}  {k}")
