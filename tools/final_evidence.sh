#!/bin/bash
# Round evidence on ONE B200 (run under gpurun): GPU tests, ncu full capture of match_kernel + traffic stamp, bench lines,
# phase cycles, launch list.  Outputs under gpurun_out/ with the given prefix.
P=${1:-r3_final}
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/${P}_gputests.log
ncu --set full --clock-control none --import-source on -k regex:match_kernel --launch-skip 19 -c 1 -f -o $O/${P}_match \
    python tools/run_steps.py c3 1024 24 > $O/${P}_ncu.log 2>&1
ncu -i $O/${P}_match.ncu-rep --page raw --csv > /tmp/raw.csv 2>/dev/null
python tools/ncu_traffic.py /tmp/raw.csv c3 1024 > $O/${P}_traffic_stamp.log 2>&1
cp profiles/r2_match_traffic.json $O/${P}_match_traffic.json
python bench.py --steps 20 --warmup 3 > $O/${P}_bench_n1.json 2> $O/${P}_bench_n1.err
python tools/phase_cycles.py c3 1024 4 0.25 > $O/${P}_phase_4.txt 2>&1
python tools/phase_cycles.py c3 1024 60 0.35 > $O/${P}_phase_60.txt 2>&1
SLAM_BENCH_MIN_TIMED_S=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > $O/${P}_launches_bench.log 2>&1
python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu > $O/${P}_bench_c2_n1.json 2> $O/${P}_bench_c2_n1.err
ncu --set full --clock-control none -k regex:update_apply_kernel --launch-skip 20 -c 1 -f -o $O/${P}_update \
    python tools/run_steps.py c3 1024 24 > $O/${P}_ncu_update.log 2>&1
python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu > $O/${P}_bench_c5_n1.json 2> $O/${P}_bench_c5_n1.err
tail -2 $O/${P}_gputests.log
python - <<PY
import json
for f in ("bench_n1", "bench_c2_n1", "bench_c5_n1"):
    try:
        d = json.loads(open("$O/${P}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), d["kernel_ms"], round(d["roofline"]["frac"], 4), d["roofline"]["traffic"])
    except Exception as e:
        print(f, "ERR", e)
PY
