"""Phase view of a barrier-structured kernel from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`: the SASS
rows are put back in address order, cut at every BAR / EXIT, and each segment's executed warp instructions and stall samples are
summed; the source lines that contributed most instructions label the segment.
    python scripts/ncu_phases.py source.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, fname, line = None, None, 0
sass = {}


def num(s):
    try:
        return int(float(s))
    except (ValueError, TypeError):
        return 0


for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":
        line = num(r[0])
        continue
    if not r[2].startswith("0x"):
        continue
    addr = int(r[2], 16)
    d = {k: v for k, v in zip(hdr[4:], r[4:])}
    sass[addr] = (r[3].strip(), num(d["Instructions Executed"]), num(d["# Samples"]),
                  {k[6:]: num(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and num(v)}, f"{fname}:{line}")
new = lambda: dict(inst=0, samp=0, n=0, ops=collections.Counter(), stalls=collections.Counter(), lines=collections.Counter())
seg, segs = new(), []
for addr in sorted(sass):
    text, ie, sm, st, ln = sass[addr]
    toks = text.split()
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    seg["inst"] += ie
    seg["samp"] += sm
    seg["n"] += 1
    seg["ops"][op.split(".")[0]] += ie
    seg["stalls"].update(st)
    seg["lines"][ln] += ie
    if op.startswith("BAR") or op.startswith("EXIT"):
        segs.append(seg)
        seg = new()
segs.append(seg)
ti = sum(s["inst"] for s in segs) or 1
ts = sum(s["samp"] for s in segs) or 1
print("total warp instructions", ti, "samples", ts)
for i, s in enumerate(segs):
    if s["inst"] / ti < 0.003 and s["samp"] / ts < 0.003:
        continue
    ops = " ".join(f"{k}:{100*v/max(s['inst'],1):.0f}" for k, v in s["ops"].most_common(7))
    st = " ".join(f"{k}:{100*v/max(s['samp'],1):.0f}" for k, v in s["stalls"].most_common(4))
    ln = " ".join(k.replace("mp_fused.cu:", "L").replace("tc_common.cuh:", "tc") for k, v in s["lines"].most_common(3))
    print(f"seg {i:2d} sass {s['n']:5d} inst {100*s['inst']/ti:5.1f}% samp {100*s['samp']/ts:5.1f}% | {ln} | {ops} | {st}")
