"""FastSLAM accuracy vs ground truth on the Intel log for several particle counts (F4):
    python tools/eval_accuracy.py [--particles 10,1024,8192] [--unit 0.05] [--frames 910] [--out gpurun_out/r2_accuracy.json]
Uses tests/golden/intel_full.npz (raw odometry + ranges + corrected-log poses of all 910 stamps), a pre-sized 72 m map,
the reference's matcher parameters (FastSlam.py:197-199) and seed 0.  The estimate is the trajectory of the particle
with the largest weight at the end of the run (FastSlam.py:165-170 picks that particle's map)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
import slam_2d_lidar_scan_b200 as S  # noqa: E402
from slam_2d_lidar_scan_b200.evaluate import evaluate_trajectory  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--particles", default="10,1024,8192")
ap.add_argument("--unit", type=float, default=0.1)
ap.add_argument("--map", type=float, default=100.0, help="fixed map side [m] (no expansion: lost particles are clipped)")
ap.add_argument("--frames", type=int, default=910)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_accuracy.json"))
a = ap.parse_args()
d = np.load(os.path.join(ROOT, "tests", "golden", "intel_full.npz"))
T = min(a.frames, len(d["poses"]))
raw, ranges, truth = d["poses"][:T], d["ranges"][:T], d["truth"][:T]
init = {"x": float(raw[0, 0]), "y": float(raw[0, 1])}
res = dict(unit=a.unit, map=a.map, frames=T, raw_odometry=evaluate_trajectory(raw, truth), runs=[])
for n in (int(v) for v in a.particles.split(",")):
    np.random.seed(0)
    pf = S.ParticleFilter(n, [a.map, a.map, init, a.unit, np.pi, 10, 180, 5 * a.unit], [1.4, 0.25, 2, 0.1, 0.25, 0.3, 0.15, 5])
    pf.keepTrajectory = False
    pf.ignoreMissingHeading = True
    pf.expandMaps = False                # a particle that leaves the 100 m map is lost anyway; its windows are clipped
    pf.ignoreStatusBits = 1 | 2 | 8
    hist = torch.zeros((T, n, 3), dtype=torch.float64, device=pf.geom.device)
    torch.cuda.synchronize()
    t0 = time.time()
    resamples = 0
    for k in range(T):
        pf.updateParticles({"x": float(raw[k, 0]), "y": float(raw[k, 1]), "theta": float(raw[k, 2]), "range": ranges[k]}, k + 1)
        if pf.weightUnbalanced():
            idxBefore = None
            pf.resample()
            resamples += 1
            hist[:k] = hist[:k].index_select(1, torch.from_numpy(pf.lastResampleIdx).to(hist.device))   # history follows the copies
        hist[k] = pf.prevMatched
    torch.cuda.synchronize()
    dt = time.time() - t0
    b = pf.best_particle()
    est = hist[:, b].cpu().numpy()
    run = dict(particles=n, seconds=round(dt, 2), particle_scans_per_s=round(n * T / dt, 1), resamples=resamples,
               best=evaluate_trajectory(est, truth))
    allAte = [evaluate_trajectory(hist[:, i].cpu().numpy(), truth)["ate"]["rmse"] for i in range(0, n, max(1, n // 64))]
    run["ate_rmse_over_particles"] = dict(min=float(np.min(allAte)), median=float(np.median(allAte)), max=float(np.max(allAte)),
                                          sampled=len(allAte))
    res["runs"].append(run)
    print(json.dumps(run), flush=True)
    del pf, hist
    torch.cuda.empty_cache()
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
print("raw odometry:", json.dumps(res["raw_odometry"]))
