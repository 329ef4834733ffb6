"""Debug helper: whole Intel log through the deterministic driver (c3 cells, 72 m map); at the first pose that differs
from the reference golden, re-run that matchScan on the oracle with the GPU's own map and report where they part."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import __graft_entry__ as g; g.build()
import slam_2d_lidar_scan_b200 as S
from oracle import slam_oracle as O
from conftest import load_golden
full, gd = load_golden("intel_full.npz"), load_golden("det_intel_full.npz")
fr = [dict(x=float(p[0]), y=float(p[1]), theta=float(p[2]), range=r.tolist()) for p, r in zip(full["poses"], full["ranges"])]
init = {"x": fr[0]["x"], "y": fr[0]["y"]}
ogA = (72, 72, init, 0.05, np.pi, 180, 10, 0.25); smA = (1.5, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 5)
og = S.OccupancyGrid(*ogA)
sm = S.ScanMatcher(og, *smA)
xT, yT = [], []
for count, f in enumerate(fr, start=1):
    cur = f
    if count == 1:
        prevRawTh = prevMatchedTh = None
        matched, conf = cur, 1
    else:
        est, dist, estTh, rawTh = S.updateEstimatedPose(cur, prevMatched, prevRaw, prevRawTh, prevMatchedTh)
        sm.debug = count >= 755
        matched, conf = sm.matchScan(est, dist, estTh, count)
        prevRawTh = rawTh
        prevMatchedTh = S.getMovingTheta(matched, xT, yT)
    if not np.array_equal([matched['x'], matched['y'], matched['theta']], gd["c3_poses"][count - 1]):
        print("MISMATCH at", count, "got", matched['x'], matched['y'], matched['theta'], "golden", gd["c3_poses"][count - 1],
              "conf", conf, gd["c3_confs"][count - 1], "idx", sm.lastIdx)
        rog = O.OccupancyGrid(*ogA)
        rog.occupancyGridVisited[:] = og.occupancyGridVisited
        rog.occupancyGridTotal[:] = og.occupancyGridTotal
        rsm = O.ScanMatcher(rog, *smA)
        rsm.trace = []
        rm, rc = rsm.matchScan(est, dist, estTh, count)
        print("oracle on the GPU map:", rm['x'], rm['y'], rm['theta'], rc, rsm.lastIdx)
        for k, stage in enumerate(("coarse", "fine")):
            a, b = sm.last[stage + "_prob"], rsm.trace[k]["prob"]
            print(stage, "prob shapes", a.shape, b.shape, "equal", a.shape == b.shape and np.array_equal(a, b))
            if a.shape == b.shape and not np.array_equal(a, b):
                d = np.argwhere(a != b); print("  prob differs at", len(d), "cells, first", d[:5], a[tuple(d[0])], b[tuple(d[0])])
            va, vb = sm.last[stage + "_vol"], rsm.trace[k]["vol"]
            print(stage, "vol equal", np.array_equal(va, vb), "argmax", np.unravel_index(va.argmax(), va.shape), np.unravel_index(vb.argmax(), vb.shape))
            if not np.array_equal(va, vb):
                d = np.argwhere(va != vb); print("  vol differs at", len(d), "first", d[:5], va[tuple(d[0])], vb[tuple(d[0])])
        break
    og.updateOccupancyGrid(matched)
    S.updateTrajectory(matched, xT, yT)
    prevMatched, prevRaw = matched, cur
print("done", count)
