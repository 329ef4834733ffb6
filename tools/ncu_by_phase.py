"""Aggregate an ncu source-page CSV (`ncu -i rep --page source --csv --print-source cuda,sass`) of match.cu by kernel
phase: warp-instructions executed, stall samples and the top stall reasons per phase.

SASS instructions inlined from other files (common.cuh's dadd/ddiv, CUDA intrinsics headers) are attributed to the
phase of the nearest preceding match.cu line in ADDRESS order, so the table follows the machine code layout.
Phase boundaries are found from marker strings in the captured source itself (no hard-coded line numbers).

    python tools/ncu_by_phase.py file.csv [particles] [source.csv from --print-source cuda of the same report]
"""
import csv
import sys
from collections import defaultdict

MARKS = [  # (marker text in match.cu, phase name) -- a phase runs from its marker to the next one
    ("struct FetchDense", "scores: fetch dense"), ("struct FetchGated", "scores: fetch gated"), ("struct PruneCtx", "scores: pairwise"),
    ("void warp_sort", "lists: sort"), ("void build_list", "lists: rotate/unique"), ("struct BlockScratch", "block reductions"),
    ("__noinline__ void blur_stage", "blur D1 dilate + tiles"), ("// D2. active tiles", "blur D2"),
    ("struct ScoreArgs", "scores: task loop/epilogue"), ("__noinline__ void lists_batch", "lists: batch loop"),
    ("void union_window", "union window geometry"), ("float2 lds_f2", "stream: pack"), ("__noinline__ void stream_role", "stream: TMA/role"),
    ("__device__ void run_stage", "A geometry"), ("// ---- B. clear", "B clear + maps"), ("// ---- C. occupied", "C scatter setup"),
    ("    if (mode == 1) {", "C scatter shift"), ("    } else if (mode == 2) {", "C scatter range"), ("    } else {\n", "C scatter generic"),
    ("    if (mode != 0) {      // transposed", "C transpose"), ("// ---- D. separable", "D/E call, min/clamp"),
    ("// ---- F. beam", "F points"), ("// ---- G. per-theta", "G driver"), ("// ---- H. select", "H select"),
    ("__global__ void __launch_bounds__", "kernel main loop + cold start"), ("__global__ void lut_kernel", "other kernels"),
]

rows = list(csv.reader(open(sys.argv[1])))
nPart = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cur = hdr = None
line = None
addr = {}          # address -> dict(file, line, samples, inst, stalls)
src = {}           # match.cu line -> text
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 2 and r[0] == "Line No":
        hdr = r
        continue
    if not hdr or len(r) < 10:
        continue
    if r[0].isdigit():
        line = int(r[0])
        if cur == "match.cu":
            src[line] = r[1]
        continue
    if r[2].startswith("0x"):
        d = dict(zip(hdr[4:], r[4:]))
        try:
            smp, inst = int(d["# Samples"]), int(d["Instructions Executed"])
        except (KeyError, ValueError):
            continue
        st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
        addr[int(r[2], 16)] = dict(file=cur, line=line, smp=smp, inst=inst, st=st, sass=r[3].strip())

if len(sys.argv) > 3 and sys.argv[3].endswith(".cu"):      # the source file itself (must be the captured version)
    for i, l in enumerate(open(sys.argv[3]).read().split("\n"), start=1):
        src[i] = l
elif len(sys.argv) > 3:      # full source text as captured in the report (lines without SASS are absent from `src`)
    for r in csv.reader(open(sys.argv[3])):
        if len(r) >= 2 and r[0].isdigit():
            src[int(r[0])] = r[1]
bounds = []
for text, name in MARKS:
    hit = [ln for ln, s in src.items() if text.strip("\n") in s and (not text.endswith("\n") or s.rstrip() == text.rstrip("\n"))]
    if text.endswith("\n") and bounds:      # a bare `} else {`: the first one AFTER the previous marker (the scatter's range path)
        hit = [ln for ln in hit if ln > bounds[-1][0]]
    if hit:
        bounds.append((min(hit), name))
bounds.sort()


def phase_of(ln):
    name = "helpers / prologue"
    for b, n in bounds:
        if ln >= b:
            name = n
    return name


agg = defaultdict(lambda: [0, 0, defaultdict(int)])
last = "helpers / prologue"
for a in sorted(addr):
    e = addr[a]
    if e["file"] == "match.cu":
        last = phase_of(e["line"])
    g = agg[last]
    g[0] += e["smp"]; g[1] += e["inst"]
    for k, v in e["st"].items():
        g[2][k] += v
ts = sum(g[0] for g in agg.values()) or 1
ti = sum(g[1] for g in agg.values()) or 1
print("total samples %d, warp-instructions %d (%.0f per particle)" % (ts, ti, ti / nPart))
for k, g in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    top = sorted(g[2].items(), key=lambda kv: -kv[1])[:5]
    print("%5.1f%% smp %5.1f%% inst %8.0f inst/particle  %-28s %s" % (100 * g[0] / ts, 100 * g[1] / ti, g[1] / nPart, k,
                                                                   " ".join("%s=%d" % t for t in top)))
