"""Dataset drivers: the loops of the reference's three ``main()`` programs on the GPU classes, without the plots
(SURVEY.md section 8f row F3).

    python -m slam_2d_lidar_scan_b200.drivers scanmatch --data DataSet/PreprocessedData/intel_gfs --map 72 --out run1
    python -m slam_2d_lidar_scan_b200.drivers fastslam  --data ... --particles 1024 --seed 0 --out run2
    python -m slam_2d_lidar_scan_b200.drivers mapping   --data ... --out run3

* ``scanmatch``  Utils/ScanMatcher_OGBased.py:226-256 (matchMax=True, deterministic)
* ``fastslam``   Algorithm/FastSlam.py:152-162 (best particle = max weight, :165-170)
* ``mapping``    Utils/OccupancyGrid.py:198-200 (known poses)

``--gt <corrected log>`` additionally scores the trajectory against ground truth (evaluate.py: ATE / RPE; the
reference's own compareGT, ScanMatcher_OGBased.py:270-289, is only a per-step debugging printout) and writes
``<out>_accuracy.json`` with the raw odometry's scores next to it.

Maps are pre-sized (``--map`` metres, centred on the first pose) because this implementation does not expand them;
the reference's defaults otherwise (unit 0.02 m, 180 beams over pi, 10 m range -- :292-294 / FastSlam.py:197-199).
Outputs ``<out>_trajectory.npy`` ([T][3] matched poses, best particle for fastslam) and ``<out>_map.npz``
(visited, total of that map).
"""
import argparse
import time

import numpy as np

from .evaluate import evaluate_trajectory
from .fastslam import ParticleFilter
from .grid import OccupancyGrid
from .matcher import ScanMatcher, getMovingTheta, readJson, updateEstimatedPose, updateTrajectory


def run_scanmatch(sensorData, og, sm, maxFrames=None, log=None):
    """Single-trajectory scan-match SLAM; returns [T][3] matched poses and the confidences."""
    xT, yT, poses, confs = [], [], [], []
    keys = sorted(sensorData.keys())[:maxFrames]
    for count, key in enumerate(keys, start=1):
        cur = sensorData[key]
        if count == 1:
            prevRawMovingTheta, prevMatchedMovingTheta = None, None
            matched, conf = cur, 1
        else:
            est, dist, estTh, rawTh = updateEstimatedPose(cur, prevMatched, prevRaw, prevRawMovingTheta,
                                                          prevMatchedMovingTheta)
            matched, conf = sm.matchScan(est, dist, estTh, count)
            prevRawMovingTheta = rawTh
            prevMatchedMovingTheta = getMovingTheta(matched, xT, yT)
        og.updateOccupancyGrid(matched)
        updateTrajectory(matched, xT, yT)
        prevMatched, prevRaw = matched, cur
        poses.append([matched['x'], matched['y'], matched['theta']])
        confs.append(conf)
        if log and count % 50 == 0:
            log("frame %d / %d" % (count, len(keys)))
    return np.array(poses), np.array(confs, dtype=np.float64)


def run_fastslam(pf, sensorData, maxFrames=None, log=None):
    """FastSLAM loop; returns the per-step pose of the currently best particle, the resample flags, the best index."""
    keys = sorted(sensorData.keys())[:maxFrames]
    best, fired = [], []
    for count, key in enumerate(keys, start=1):
        pf.updateParticles(sensorData[key], count)
        f = pf.weightUnbalanced()
        if f:
            pf.resample()
        fired.append(f)
        b = pf.best_particle()
        best.append(pf.prevMatched[b].cpu().numpy())
        if log and count % 50 == 0:
            log("frame %d / %d  variance %.3g" % (count, len(keys), pf.lastVariance))
    return np.array(best), np.array(fired), pf.best_particle()


def run_mapping(sensorData, og, maxFrames=None):
    keys = sorted(sensorData.keys())[:maxFrames]
    for key in keys:
        og.updateOccupancyGrid(sensorData[key])
    return np.array([[sensorData[k]['x'], sensorData[k]['y'], sensorData[k]['theta']] for k in keys])


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("mode", choices=["scanmatch", "fastslam", "mapping"])
    ap.add_argument("--data", required=True, help="preprocessed JSON ({'map': {timestamp: {x, y, theta, range}}})")
    ap.add_argument("--out", default="slam_run")
    ap.add_argument("--map", type=float, default=72.0, help="pre-sized square map side in metres")
    ap.add_argument("--unit", type=float, default=0.02)
    ap.add_argument("--particles", type=int, default=10)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--device", default=None)
    ap.add_argument("--gt", default=None, help="ground-truth JSON of the same stamps (e.g. intel_corrected_log)")
    a = ap.parse_args(argv)
    data = readJson(a.data)
    first = data[sorted(data.keys())[0]]
    K = len(first['range'])
    unit, fov, maxRange = a.unit, np.pi, 10
    initXY = {"x": first['x'], "y": first['y']}
    smArgs = [1.4, 0.25, 2, 0.1, 0.25, 0.3, 0.15, 5]
    t0 = time.time()
    log = lambda m: print("[%.1fs] %s" % (time.time() - t0, m), flush=True)
    if a.mode == "fastslam":
        if a.seed is not None:
            np.random.seed(a.seed)
        pf = ParticleFilter(a.particles, [a.map, a.map, initXY, unit, fov, maxRange, K, 5 * unit], smArgs, device=a.device)
        traj, fired, b = run_fastslam(pf, data, a.frames, log)
        og = pf.particles[b].og
        log("resamples: %d" % int(fired.sum()))
    else:
        wall = (7 if a.mode == "mapping" else 5) * unit
        og = OccupancyGrid(a.map, a.map, initXY, unit, fov, K, maxRange, wall, device=a.device)
        if a.mode == "mapping":
            traj = run_mapping(data, og, a.frames)
        else:
            sm = ScanMatcher(og, *smArgs)
            traj, _ = run_scanmatch(data, og, sm, a.frames, log)
    np.save(a.out + "_trajectory.npy", traj)
    np.savez_compressed(a.out + "_map.npz", visited=og.occupancyGridVisited.astype(np.float32),
                        total=og.occupancyGridTotal.astype(np.float32), mapXLim=og.mapXLim, mapYLim=og.mapYLim)
    log("%d frames -> %s_trajectory.npy, %s_map.npz" % (len(traj), a.out, a.out))
    if a.gt:
        import json
        gt = readJson(a.gt)
        keys = sorted(data.keys())[:len(traj)]
        truth = np.array([[gt[k]['x'], gt[k]['y'], gt[k]['theta']] for k in keys])
        raw = np.array([[data[k]['x'], data[k]['y'], data[k]['theta']] for k in keys])
        acc = dict(mode=a.mode, particles=a.particles if a.mode == "fastslam" else 1, unit=unit,
                   estimate=evaluate_trajectory(traj, truth), raw_odometry=evaluate_trajectory(raw, truth))
        json.dump(acc, open(a.out + "_accuracy.json", "w"), indent=1)
        log("ATE rmse %.3f m (raw odometry %.3f m), RPE %.4f m / %.4f rad per step" % (
            acc["estimate"]["ate"]["rmse"], acc["raw_odometry"]["ate"]["rmse"], acc["estimate"]["rpe"]["trans_rmse"],
            acc["estimate"]["rpe"]["rot_rmse"]))


if __name__ == "__main__":
    main()
