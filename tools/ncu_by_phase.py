"""Aggregate an ncu source-page CSV of match.cu by kernel phase (line ranges): samples, instructions, top stalls.
python tools/ncu_by_phase.py file.csv"""
import csv
import sys
from collections import defaultdict

RANGES = [  # (first line, last line, phase) in csrc/match.cu -- keep in sync when the file moves
    (174, 335, "scores: gathers + pairwise sums"), (342, 431, "lists: sort/unique"), (440, 463, "block reductions"),
    (480, 737, "blur (D1 dilate + D2)"), (750, 803, "scores: task loop/epilogue"), (811, 831, "lists: batch loop"),
    (832, 846, "union window geometry"), (854, 990, "stream warps (TMA + pack)"), (1004, 1065, "A/B geometry, clear, maps"),
    (1066, 1267, "C scatter + transpose"), (1268, 1308, "D/E blur call, min/clamp"), (1309, 1349, "F points"),
    (1350, 1384, "G lists+scores driver"), (1385, 1506, "H select"), (1507, 1561, "kernel main loop"),
    (114, 173, "helpers (lds, mbarrier wait, tma)"), (44, 61, "csync"),
]
rows = list(csv.reader(open(sys.argv[1])))
cur = hdr = None
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1]; continue
    if len(r) > 2 and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) >= 10 and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            smp, inst = int(d["# Samples"]), int(d["Instructions Executed"])
        except (KeyError, ValueError):
            continue
        fn = cur.split('/')[-1]
        ln = int(r[0])
        key = fn
        if fn == "match.cu":
            key = "match.cu:other"
            for a, b, name in RANGES:
                if a <= ln <= b:
                    key = name; break
        e = agg[key]
        e[0] += smp; e[1] += inst
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v.isdigit():
                e[2][k[6:]] += int(v)
ts = sum(e[0] for e in agg.values()) or 1
ti = sum(e[1] for e in agg.values()) or 1
print("total samples %d, warp-instructions %d" % (ts, ti))
for k, e in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    top = sorted(e[2].items(), key=lambda kv: -kv[1])[:4]
    print("%5.1f%% smp %5.1f%% inst  %-36s %s" % (100 * e[0] / ts, 100 * e[1] / ti, k, " ".join("%s=%d" % t for t in top)))
