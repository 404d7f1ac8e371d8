"""Per-source-line instruction and stall-sample shares from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
    python scripts/ncu_lines.py source.csv [min_pct]"""
import collections
import csv
import sys


def num(s):
    try:
        return int(float(s))
    except (ValueError, TypeError):
        return 0


rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
hdr, mode = None, None
agg = collections.defaultdict(lambda: [0, 0])
stall = collections.defaultdict(collections.Counter)
fname = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1]
    if len(r) > 5 and r[0] == "Line No":
        hdr, mode = r, "cuda"
        continue
    if len(r) > 5 and r[0] == "Address":
        hdr, mode = r, "sass"
        continue
    if hdr is None or len(r) < len(hdr) or mode != "cuda":
        continue
    d = dict(zip(hdr, r))
    ln = (fname.split("/")[-1], num(d["Line No"]))
    agg[ln][0] += num(d.get("Instructions Executed"))
    agg[ln][1] += num(d.get("# Samples"))
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k and num(d[k]):
            stall[ln][k] += num(d[k])
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print("instructions", ti, "samples", ts)
for ln in sorted(agg):
    i, s = agg[ln]
    if 100 * i / ti > thr or 100 * s / ts > thr:
        top = ", ".join(f"{k[6:]}:{v}" for k, v in stall[ln].most_common(3))
        print(f"{ln[0]}:{ln[1]:4d} inst {100*i/ti:5.1f}% samp {100*s/ts:5.1f}%  {top}")
