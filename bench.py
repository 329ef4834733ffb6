#!/usr/bin/env python
"""Benchmark of the scan-match / FastSLAM hot path (BASELINE.json metric: particle-scans/s + HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c2|c5]

A "step" is one FastSLAM step over the rank's particles: odometry proposal, priors, fused coarse+fine scan match,
heading/weight update, map update, weight normalisation (+ the all-gather when N > 1).  Weak scaling: every GPU
holds the workload's particle count (c3: 1024 particles, 1001x1001 lattices @0.05 m, 180 beams, 10 440 poses).

  value   whole-job particle-scans/s with the step's inputs already resident in HBM (pre-staged), CUDA-event timed
  e2e     the same through the public API (ParticleFilter.updateParticles + weightUnbalanced) with HOST readings:
          per step one pinned H2D copy (ranges | uniforms | radial prior) and a D2H read of (variance, trigger, status).
          In both arms the normalisation / trigger (/ all-gather) of a step run on a side stream next to its map update
          (engine.SideTrigger); the timed regions end with a full device synchronisation
  roofline  fused match kernel: N * B_match algorithmic bytes / its mean CUDA-event duration, vs the measured HBM peak;
          `traffic` comes from profiles/r2_match_traffic.json only while that file's hash of csrc/match.cu + common.cuh
          still matches the sources (tools/ncu_traffic.py stamps it), else null
  timing  the K-step schedule is replayed until the timed region is >= ~1 s (`repeats`, `timed_steps`; a replay
          restarts from the same particle poses, the lattices keep integrating), so that the clock sampler sees tens
          of loaded samples; the synthetic odometry steps 0.35 m, so the reference's heading
          prior (mode 1, priors_kernel) is part of every timed step; `kernel_ms` = mean CUDA-event duration per launch
  cpu_baseline  the oracle port of the reference's numpy path on the host cores (rank 0, N=1, bounded sample)

--impl reference times the oracle port (the reference is pure Python and is not present on the GPU box) on all
host cores for the same workload/metric.
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle-scans/s (180-beam scan matched + mapped per particle)"
STRIDE = 0.35          # synthetic odometry step [m]: > 0.3 m, so the heading prior is active (FastSlam.py:88)
MIN_TIMED_S = float(os.environ.get("SLAM_BENCH_MIN_TIMED_S", "1.0"))   # the timed regions run at least this long
                       # (the variable exists for the ncu launch-list pass only: a profiler serialises every launch)


# ------------------------------------------------------------------------------------------------ CPU baseline
def _cpu_worker(args):
    """One process = one particle of the oracle (the reference's numpy path restated), timed over `steps` steps."""
    workload, steps, seed = args
    from oracle import slam_oracle as O
    spec = importlib_pkg().synthetic.config(workload)
    scene = importlib_pkg().synthetic.make_scene(seed=0, steps=steps + 1, K=spec["K"], fov=spec["og"][4],
                                                 unit=spec["og"][3], stride=STRIDE)
    np.random.seed(seed)
    p = O.Particle(spec["og"], spec["sm"])
    for fr in scene["warm"]:
        p.og.updateOccupancyGrid(fr)
    p.update(scene["frames"][0], 1)
    t0 = time.perf_counter()
    for count, fr in enumerate(scene["frames"][1:steps + 1], start=2):
        p.update(fr, count)
    return time.perf_counter() - t0


def importlib_pkg():
    import importlib.util
    name = "slam_b200_synthetic_only"
    if name in sys.modules:
        return sys.modules[name]
    # the synthetic scene generator is numpy-only; load it without importing the CUDA package

    class _Pkg:
        pass
    spec = importlib.util.spec_from_file_location("slam_b200_synth", os.path.join(ROOT, "slam-2d-lidar-scan_b200",
                                                                                   "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    pkg = _Pkg()
    pkg.synthetic = mod
    sys.modules[name] = pkg
    return pkg


def host_cores():
    """CPUs this process may actually use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except (OSError, ValueError):
        pass
    return n


def cpu_baseline(workload, steps=6, procs=None):
    """Oracle port on every usable host core: persistent worker pool (one particle per process), one untimed warm-up
    round (first-call import / allocation costs), then `steps` timed scans per process."""
    procs = procs or host_cores()
    t0 = time.perf_counter()
    rate, times = oracle_pool_rate(workload, per=steps, rounds=1, warm=1, procs=procs)
    wall = time.perf_counter() - t0
    return dict(value=rate, unit="particle-scans/s", cores=procs, kind="port",
                sample="%d processes x 1 particle x %d steps of workload %s after 1 warm-up round (oracle numpy port; "
                       "%.1f s wall incl. setup)" % (procs, steps, workload, wall)), times[0]


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, index, enabled=True):
        self.index, self.rows, self.proc, self.enabled = index, [], None, enabled

    def __enter__(self):
        if not self.enabled:      # only rank 0 polls nvidia-smi: N pollers would steal host cores from the N ranks
            return self
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def kernel_source_hash():
    """sha256 over the sources of the match kernel: a traffic capture is only valid for the code it was taken from."""
    import hashlib
    h = hashlib.sha256()
    for f in ("match.cu", "common.cuh"):
        h.update(open(os.path.join(ROOT, "slam-2d-lidar-scan_b200", "csrc", f), "rb").read())
    return h.hexdigest()


def measured_traffic(workload, nLocal):
    """dram__bytes_read.sum + dram__bytes_write.sum of one match_kernel launch from the committed `ncu --set full`
    capture (profiles/r2_match_traffic.json, written by tools/ncu_traffic.py), or None when the capture is of another
    workload or of other kernel sources (stale)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_match_traffic.json")))
    except (OSError, ValueError):
        return None
    if t.get("workload") == workload and t.get("particles") == nLocal and t.get("kernel_sha256") == kernel_source_hash():
        return t.get("dram_bytes_per_launch")
    return None


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    import slam_2d_lidar_scan_b200 as S
    from slam_2d_lidar_scan_b200 import synthetic
    from slam_2d_lidar_scan_b200.distributed import ShardedParticleFilter

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec = synthetic.config(args.workload)
    nLocal = args.particles or spec["N"]
    K, W = args.steps, args.warmup
    np.random.seed(1234)                                  # same stream on every rank (sharding is invisible)
    spf = ShardedParticleFilter(nLocal * world, spec["og"], spec["sm"], device=dev)
    pf = spf.local
    pf.keepTrajectory = False
    pf.expandMaps = False               # pre-sized 50 m lattices; a window leaving them fails the run
    pf.ignoreMissingHeading = True      # ~1e5 sampled particle-steps: the reference's None + float TypeError (a particle that
                                        # did not move before a > 0.3 m odometry step) does occur; keep going like the kernels do
    geomBytes = pf.grids.numel() * 4

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def fresh_scene(steps):
        """Maps pre-warmed with 8 scans at true poses, identical for all particles; first reading consumed."""
        scene = synthetic.make_scene(seed=0, steps=steps + 2, K=spec["K"], fov=spec["og"][4], unit=spec["og"][3],
                                     stride=STRIDE)
        return scene["frames"]

    og = S.OccupancyGrid(*pf.geom.args, _geometry=pf.geom)
    for fr in synthetic.make_scene(seed=0, steps=1, K=spec["K"], fov=spec["og"][4], unit=spec["og"][3],
                                   stride=STRIDE)["warm"]:
        og.updateOccupancyGrid(fr)
    pf.grids.copy_(og.device_grid.unsqueeze(0).expand_as(pf.grids))
    del og

    # ---- pilot: a few steps through the public API to size the timed regions (>= MIN_TIMED_S each)
    frames = fresh_scene(8 + 3 * (W + K))
    count = 0
    for fr in frames[:3]:
        count += 1
        spf.updateParticles(fr, count)
        spf.weightUnbalanced()
    sync_all()
    t0 = time.perf_counter()
    for fr in frames[3:6]:
        count += 1
        spf.updateParticles(fr, count)
        spf.weightUnbalanced()
    sync_all()
    pilot = torch.tensor([(time.perf_counter() - t0) / 3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(pilot, op=dist.ReduceOp.MAX)
    repeats = max(1, int(np.ceil(MIN_TIMED_S / (float(pilot.item()) * K))))
    KT = K * repeats                                      # timed steps per region

    # The K-step schedule is replayed `repeats` times.  A replay starts from the same particle poses / headings /
    # weights (the lattices keep integrating the scans): the reference never resamples a power-of-two population
    # (SURVEY A9), so hundreds of consecutive sampled steps would only measure a diverging random walk.
    def snapshot():
        spf.flush()                      # a deferred copy-back of normalised weights lands before they are saved
        return (pf.prevMatched.clone(), pf.prevHeading.clone(), pf.hasHeading.clone(), pf.weights.clone(),
                list(pf._prevRaw), list(pf._prevRawHeading))

    def restore(sn):
        spf.flush()
        pf.prevMatched.copy_(sn[0]); pf.prevHeading.copy_(sn[1]); pf.hasHeading.copy_(sn[2]); pf.weights.copy_(sn[3])
        pf._prevRaw, pf._prevRawHeading = list(sn[4]), list(sn[5])

    # ---- (1) inputs resident in HBM: pre-stage every step's [ranges | uniforms | rv] row
    recs, rows = [], []
    prevRaw, prevHead = pf._prevRaw[0], pf._prevRawHeading[0]
    for i in range(W + K):
        count += 1
        fr = frames[count - 1]
        u = np.random.random_sample(nLocal * world)[spf.lo:spf.hi]
        row = torch.zeros_like(pf._stage_h)
        rec = pf._prepare(fr, count, nLocal, prevRaw, prevHead, out=row, uniforms=u)
        recs.append(rec); rows.append(row.to(dev))
        prevRaw, prevHead = fr, rec["newRawHeading"]
    mode1 = sum(1 for r in recs[W:] if r["mode"] == 1) * repeats
    for i in range(W):
        pf._launch(0, nLocal, recs[i], rows[i])
        spf.gather_and_normalize()
    snap = snapshot()
    sync_all()
    with ClockSampler(local, rank == 0) as clk1:
        pf.matchEvents = []
        launches0 = pf.kernelLaunches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(dev))
        for r in range(repeats):
            restore(snap)
            for i in range(W, W + K):
                pf._launch(0, nLocal, recs[i], rows[i])
                spf.gather_and_normalize()
        spf.flush()                      # the last step's side-stream work is inside the timed region
        e1.record(torch.cuda.current_stream(dev))
        sync_all()
    pf._prevRaw, pf._prevRawHeading = [prevRaw] * nLocal, [prevHead] * nLocal
    launches = (pf.kernelLaunches - launches0) // repeats            # per K steps
    ms_dev = e0.elapsed_time(e1)
    matchMs = [a.elapsed_time(b) for a, b in pf.matchEvents]
    pf.matchEvents = None
    st = int(torch.bitwise_and(pf.status, ~16).max().item())      # sticky OR-ed status words (HEADING_MISSING tolerated, see above)
    if st:
        raise RuntimeError("status bits %d set during the timed run" % st)

    # ---- (2) end to end through the public API with host readings
    def api_step(fr):
        nonlocal count
        count += 1
        spf.updateParticles(fr, count)
        return spf.weightUnbalanced()
    base = count
    for fr in frames[base:base + W]:
        api_step(fr)
    snap = snapshot()
    sync_all()
    h0, d0 = pf.h2dBytes, pf.d2hBytes
    with ClockSampler(local, rank == 0) as clk2:
        t0 = time.perf_counter()
        for r in range(repeats):
            restore(snap)
            for fr in frames[base + W:base + W + K]:
                api_step(fr)
        sync_all()
        t_e2e = time.perf_counter() - t0
    h2d, d2h = (pf.h2dBytes - h0) // KT, (pf.d2hBytes - d0) // KT

    # ---- (3) per-kernel breakdown: a few more steps with CUDA events around every launch (not part of the timings)
    pf.timer.enabled = True
    restore(snap)
    for fr in frames[base + W:base + W + min(K, 8)]:
        api_step(fr)
    kernel_ms = pf.timer.mean_ms()
    pf.timer.enabled = False

    sharding = None
    if world > 1:
        from slam_2d_lidar_scan_b200.distributed import sharding_self_check
        sharding = "pass" if sharding_self_check(dev) else "fail"

    # max over ranks
    t = torch.tensor([ms_dev, t_e2e * 1e3, statistics.mean(matchMs)], dtype=torch.float64, device=dev)
    K = KT                                                # everything below is per timed step
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_match = (float(v) for v in t.cpu())
    if rank == 0:
        total = nLocal * world
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        eng = pf.engine
        Wf = int(2 * eng.windowRadius / pf.geom.unitGridSize) + 1
        bmatch = 2 * Wf * Wf * 8                       # window read once per stage, 8 B per cell
        achieved = nLocal * bmatch / (ms_match * 1e-3) / 1e9
        c1, c2 = clk1.summary(), clk2.summary()
        out = {
            "impl": "b200", "metric": METRIC, "value": total * K / (ms_dev * 1e-3), "unit": "particle-scans/s",
            "n_gpus": world, "steps": args.steps, "warmup": W, "repeats": repeats, "timed_steps": KT, "ms_per_step": ms_dev / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "particles_per_gpu": nLocal, "particles_total": total,
                       "beams": spec["K"], "grid": "%dx%d @%.2f m" % (pf.geom.G, pf.geom.G, pf.geom.unitGridSize),
                       "poses_per_scan": sum(int(np.prod(eng.volume_shape(s))) for s in (0, 1)),
                       "parallelism": "particles sharded, %d/GPU" % nLocal,
                       "l2": "inputs larger than L2 (%.1f GB of lattices per GPU vs 126 MB L2)" % (geomBytes / 1e9),
                       "scene": "synthetic room seed 0 (SURVEY 8d), maps pre-warmed with 8 scans, odometry step "
                                "%.2f m (heading prior active in %d of %d timed steps)" % (STRIDE, mode1, KT)},
            "e2e": {"value": total * K / (ms_e2e * 1e-3), "unit": "particle-scans/s", "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "kernel_ms": {k: round(v, 5) for k, v in sorted(kernel_ms.items())},
            "roofline": {"bound": "hbm", "kernel": "slam::match_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(args.workload, nLocal),
                         "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6650",
                         "algorithmic_bytes_per_launch": nLocal * bmatch, "ms_per_launch": ms_match,
                         "share_of_step": ms_match / (ms_dev / K)},
            "clocks": {"sm_mhz": c1["sm_mhz"], "sm_max_mhz": c1["sm_max_mhz"],
                       "reasons": sorted(set(c1["reasons"]) | set(c2["reasons"])), "e2e_sm_mhz": c2["sm_mhz"],
                       "samples": c1.get("samples")},
        }
        if sharding is not None:
            out["sharding_check"] = sharding
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"], _ = cpu_baseline(args.workload, steps=args.cpu_steps)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ reference arm
_REF = {}


def _ref_init(workload, seed_base):
    """Pool initializer: every worker process owns one oracle particle, maps pre-warmed, first reading consumed."""
    from oracle import slam_oracle as O
    spec = importlib_pkg().synthetic.config(workload)
    scene = importlib_pkg().synthetic.make_scene(seed=0, steps=70, K=spec["K"], fov=spec["og"][4], unit=spec["og"][3],
                                                 stride=STRIDE)
    np.random.seed(seed_base + os.getpid() % 1000)
    p = O.Particle(spec["og"], spec["sm"])
    for fr in scene["warm"]:
        p.og.updateOccupancyGrid(fr)
    p.update(scene["frames"][0], 1)
    _REF.update(particle=p, frames=scene["frames"], count=1)


def _ref_advance(n):
    """Advance this worker's particle by n scan-match + map-update steps; returns the time spent."""
    p, frames = _REF["particle"], _REF["frames"]
    t0 = time.perf_counter()
    for _ in range(n):
        _REF["count"] += 1
        c = _REF["count"]
        p.update(frames[(c - 1) % len(frames)] if c <= len(frames) else frames[-1], c)
    return time.perf_counter() - t0


def oracle_pool_rate(workload, per, rounds, warm, procs):
    """(particle-scans/s, per-round wall times) of `procs` oracle particles advancing `per` scans per round."""
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs, initializer=_ref_init, initargs=(workload, 100)) as pool:
        pool.map(_ref_advance, [0] * procs)               # all workers initialised
        times = []
        for i in range(warm + rounds):
            t0 = time.perf_counter()
            pool.map(_ref_advance, [per] * procs, chunksize=1)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    return procs * per * rounds / sum(times), times


def run_reference(args):
    """The reference's CPU implementation of the path (oracle numpy port: the reference is pure Python and is not on
    the GPU box) on all usable host cores.  One bench step = every core advancing its own particle by `per` steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = args.steps, args.warmup
    procs = host_cores()
    spec = importlib_pkg().synthetic.config(args.workload)
    per = 2 if (K + W) <= 25 else 1                      # bounded sample: keep the whole run to a couple of minutes
    per = min(per, max(1, 60 // max(K + W, 1)))
    _, times = oracle_pool_rate(args.workload, per=per, rounds=K, warm=W, procs=procs)
    total_t = sum(times)
    value = procs * per * K / total_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "particle-scans/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": total_t / K * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "beams": spec["K"], "note": "oracle numpy port of the reference's CPU path; "
                   "the reference is pure Python and absent on the GPU box"},
        "cpu_baseline": {"value": value, "unit": "particle-scans/s", "cores": procs, "kind": "port",
                         "sample": "%d processes x 1 particle x %d scan(s) per bench step" % (procs, per)},
        "e2e": {"value": value, "unit": "particle-scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c2", "c3", "c5"])
    ap.add_argument("--particles", type=int, default=0, help="particles per GPU (default: the workload's)")
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and "RANK" not in os.environ:        # plain `python bench.py --gpus N`: start the N ranks ourselves
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if "RANK" in os.environ and world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE is %d" % (args.gpus, world))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
