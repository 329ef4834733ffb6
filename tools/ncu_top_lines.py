"""Top source lines of an ncu source-page CSV (cuda,sass view): python tools/ncu_top_lines.py file.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out, cur, hdr = [], None, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1]
        continue
    if len(r) > 2 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) >= 10 and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            samples = int(d["# Samples"]); inst = int(d["Instructions Executed"])
        except (KeyError, ValueError):
            continue
        stalls = {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        out.append((samples, inst, cur.split('/')[-1], int(r[0]), r[1].strip()[:90], top))
tot = sum(o[0] for o in out) or 1
toti = sum(o[1] for o in out) or 1
print("total samples", tot, "total warp-inst", toti)
for o in sorted(out, reverse=True)[:n]:
    print("%5.1f%% smp %5.1f%% inst %s:%d  %s  %s" % (100 * o[0] / tot, 100 * o[1] / toti, o[2], o[3], o[4],
                                                     " ".join("%s=%d" % (k[6:], v) for k, v in o[5])))
