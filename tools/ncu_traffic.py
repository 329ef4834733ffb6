"""profiles/r2_match_traffic.json from an `ncu --set full` capture of match_kernel (raw page CSV), stamped with the hash
of the kernel sources it was taken from (bench.py prints `traffic: null` when the hash no longer matches):
    ncu -i gpurun_out/r2_match.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_traffic.py /tmp/raw.csv c3 1024 [out.json]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402

rows = list(csv.reader(open(sys.argv[1])))
h, u, v = rows[0], rows[1], rows[2]
col = {k: i for i, k in enumerate(h)}


def val(name):
    if name not in col:
        hits = [k for k in col if k.endswith(name)]
        if not hits:
            return None
        name = hits[0]
    if v[col[name]] in ("", "n/a"):
        return None
    x, unit = float(v[col[name]].replace(",", "")), u[col[name]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3, "%": 1, "": 1,
             "inst": 1, "cycle": 1}
    return x * scale.get(unit, 1)


out = {
    "workload": sys.argv[2], "particles": int(sys.argv[3]), "kernel_sha256": kernel_source_hash(),
    "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
    "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
    "gpu_time_ms_under_ncu": val("gpu__time_duration.sum"),
    "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "dram_throughput_pct": val("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    "fp64_pipe_pct": val("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "lsu_shared_wavefronts": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "lsu_shared_bank_conflicts": val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    "inst_executed": val("smsp__inst_executed.sum"),
    "source": "ncu --set full --clock-control none, one launch of slam::match_kernel (tools/run_steps.py), see profiles/r2_findings.md",
}
dst = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "r2_match_traffic.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out))
