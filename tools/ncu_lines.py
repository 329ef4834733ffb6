"""Per-source-line view of an ncu `--print-source cuda,sass` CSV for a line range of match.cu:
python tools/ncu_lines.py file.csv first last [particles]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
lo, hi = int(sys.argv[2]), int(sys.argv[3])
nPart = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
cur = hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]; continue
    if len(r) > 2 and r[0] == "Line No":
        hdr = r; continue
    if hdr and cur == "match.cu" and len(r) > 10 and r[0].isdigit() and lo <= int(r[0]) <= hi:
        d = dict(zip(hdr[4:], r[4:]))
        try:
            smp, inst = int(d["# Samples"]), int(d["Instructions Executed"])
        except (KeyError, ValueError):
            continue
        st = sorted(((k[6:], int(v)) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)), key=lambda kv: -kv[1])[:3]
        print("%5d %7d smp %8.0f inst/p  %-100s %s" % (int(r[0]), smp, inst / nPart, r[1].strip()[:100], " ".join("%s=%d" % t for t in st)))
