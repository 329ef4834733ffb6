"""Generate the golden fixtures by running the UNMODIFIED reference in the build container.

Run once (CPU, needs /root/reference; matplotlib is stubbed because the image lacks it):

    python tests/golden/make_golden.py

Writes (committed, small):
    tests/golden/intel_gfs_head.json   first 40 readings of DataSet/PreprocessedData/intel_gfs (input only)
    tests/golden/det_c3.npz            deterministic matchMax=True driver, c3 geometry (unit 0.05, 50 m map), 30 frames
    tests/golden/det_ref02.npz         same driver, reference defaults at unit 0.02 on a 50 m map, 12 frames
    tests/golden/pf_c3.npz             seeded FastSLAM (np.random.seed(0)), 3 particles, c3 geometry, 22 frames
    tests/golden/update_c3.npz         mapping with known poses (OccupancyGrid.updateOccupancyGrid only)
    tests/golden/resample.npz          ParticleFilter.resample / weightUnbalanced known answers
    tests/golden/csail_gfs_head.json   first 20 readings of DataSet/PreprocessedData/csail_gfs (361 beams; input only)
    tests/golden/det_csail.npz         deterministic driver on the CSAIL readings, c3-like geometry, 20 frames
  with --full (minutes of CPU):
    tests/golden/intel_full.npz        all 910 readings of intel_gfs (raw odometry poses + ranges) and the poses of
                                       intel_corrected_log (ground truth, same keys) as float64 arrays (input only)
    tests/golden/det_intel_full.npz    deterministic driver over all 910 readings on pre-sized 72 m maps: reference
                                       defaults at unit 0.02 (BASELINE.md section 2 trajectory) and c3 parameters at 0.05

Nothing here is product code; the reference is only imported, never copied.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SLAM_REF", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_mplstub"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "Algorithm"))

from Utils.OccupancyGrid import OccupancyGrid            # noqa: E402
from Utils import ScanMatcher_OGBased as smmod            # noqa: E402
import FastSlam as fsmod                                  # noqa: E402

fsmod.print = lambda *a, **k: None
smmod.print = lambda *a, **k: None

N_FRAMES = 40


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_frames(name="intel_gfs", n=N_FRAMES):
    with open(os.path.join(REF, "DataSet/PreprocessedData", name)) as f:
        data = json.load(f)["map"]
    keys = sorted(data.keys())[:n]
    return [dict(key=k, x=data[k]["x"], y=data[k]["y"], theta=data[k]["theta"], range=data[k]["range"]) for k in keys]


def sparse_counts(og):
    """Cells that differ from the initial (1, 2) counts: flat index, visited, total."""
    v, t = og.occupancyGridVisited, og.occupancyGridTotal
    idx = np.flatnonzero((v != 1) | (t != 2))
    return idx.astype(np.int32), v.reshape(-1)[idx].astype(np.int16), t.reshape(-1)[idx].astype(np.int16)


def run_deterministic(frames, ogArgs, smArgs, nFrames, traceSteps):
    """The loop of Utils/ScanMatcher_OGBased.py:226-256 (matchMax defaults True)."""
    og = OccupancyGrid(*ogArgs)
    sm = smmod.ScanMatcher(og, *smArgs)
    trace = {}
    orig = sm.searchToMatch
    stage = {"n": 0, "count": 0}

    def spy(probSP, *a, **k):
        out = orig(probSP, *a, **k)
        c = stage["count"]
        if c in traceSteps:
            tag = "c%d_%s" % (c, "fine" if k.get("fineSearch") else "coarse")
            trace[tag + "_vol"] = out[3].copy()
            trace[tag + "_prob_sha"] = np.frombuffer(bytes.fromhex(sha(probSP)), dtype=np.uint8)
            trace[tag + "_prob_shape"] = np.array(probSP.shape)
            trace[tag + "_prob_sum"] = np.array([probSP.sum(), probSP.min(), float((probSP == 0).sum())])
        return out
    sm.searchToMatch = spy
    poses, confs = [], []
    xT, yT = [], []
    shape0 = og.occupancyGridVisited.shape
    for count, rd in enumerate(frames[:nFrames], start=1):
        stage["count"] = count
        cur = {"x": rd["x"], "y": rd["y"], "theta": rd["theta"], "range": rd["range"]}
        if count == 1:
            prevRawMovingTheta, prevMatchedMovingTheta = None, None
            matched, conf = cur, 1
        else:
            est, dist, estTh, rawTh = smmod.updateEstimatedPose(cur, prevMatched, prevRaw, prevRawMovingTheta,
                                                                prevMatchedMovingTheta)
            matched, conf = sm.matchScan(est, dist, estTh, count)
            prevRawMovingTheta = rawTh
            prevMatchedMovingTheta = smmod.getMovingTheta(matched, xT, yT)
        og.updateOccupancyGrid(matched)
        smmod.updateTrajectory(matched, xT, yT)
        prevMatched, prevRaw = matched, cur
        poses.append([matched["x"], matched["y"], matched["theta"]])
        confs.append(conf)
        assert og.occupancyGridVisited.shape == shape0, "map expanded: fixture would leave the parity regime"
    idx, v, t = sparse_counts(og)
    return dict(poses=np.array(poses), confs=np.array(confs, dtype=np.float64), cells=idx, visited=v, total=t,
                G=np.array(shape0), **trace)


def run_fastslam(frames, ogParams, smParams, nParticles, nFrames, seed):
    """The loop of Algorithm/FastSlam.py:152-162 (plots skipped)."""
    np.random.seed(seed)
    pf = fsmod.ParticleFilter(nParticles, ogParams, smParams)
    shape0 = pf.particles[0].og.occupancyGridVisited.shape
    poses, weights, resampled = [], [], []
    for count, rd in enumerate(frames[:nFrames], start=1):
        cur = {"x": rd["x"], "y": rd["y"], "theta": rd["theta"], "range": rd["range"]}
        pf.updateParticles(cur, count)
        raw = [p.weight for p in pf.particles]
        fired = pf.weightUnbalanced()
        if fired:
            pf.resample()
        resampled.append(fired)
        poses.append([[p.prevMatchedReading["x"], p.prevMatchedReading["y"], p.prevMatchedReading["theta"]]
                      for p in pf.particles])
        weights.append([raw, [p.weight for p in pf.particles]])
        for p in pf.particles:
            assert p.og.occupancyGridVisited.shape == shape0, "map expanded"
    out = dict(poses=np.array(poses), weights=np.array(weights, dtype=np.float64), resampled=np.array(resampled),
               next_uniform=np.array([np.random.random_sample()]), G=np.array(shape0))
    for i, p in enumerate(pf.particles):
        idx, v, t = sparse_counts(p.og)
        out["cells%d" % i], out["visited%d" % i], out["total%d" % i] = idx, v, t
    return out


def run_update_only(frames, ogArgs, n):
    """Utils/OccupancyGrid.py:198-200: mapping with the (raw) poses; plus the lattice/sector-table pins."""
    og = OccupancyGrid(*ogArgs)
    for rd in frames[:n]:
        og.updateOccupancyGrid({"x": rd["x"], "y": rd["y"], "theta": rd["theta"], "range": rd["range"]})
    idx, v, t = sparse_counts(og)
    L = len(og.radByX)
    sizes = np.array([len(a) for a in og.radByX])
    return dict(cells=idx, visited=v, total=t, G=np.array(og.occupancyGridVisited.shape),
                gridX_sha=np.frombuffer(bytes.fromhex(sha(og.OccupancyGridX[0])), dtype=np.uint8),
                gridY_sha=np.frombuffer(bytes.fromhex(sha(og.OccupancyGridY[:, 0])), dtype=np.uint8),
                spoke_sizes=sizes, numSpokes=np.array([og.numSpokes, og.spokesStartIdx, L]),
                spoke7_r_sha=np.frombuffer(bytes.fromhex(sha(og.radByR[7])), dtype=np.uint8),
                spoke200_x_sha=np.frombuffer(bytes.fromhex(sha(og.radByX[200])), dtype=np.uint8))


def run_resample():
    """Known answers for ParticleFilter.normalizeWeights / weightUnbalanced / resample (FastSlam.py:30-62)."""
    class P:      # stand-in with the two attributes the filter touches
        def __init__(self, w):
            self.weight = w
    out = {}
    rng = np.random.RandomState(7)
    for n in (5, 10, 15, 64):
        pf = fsmod.ParticleFilter.__new__(fsmod.ParticleFilter)
        pf.numParticles = n
        w0 = rng.random_sample(n) ** 8 * 10.0 ** rng.randint(-20, 1, n)
        pf.particles = [P(float(w)) for w in w0]
        fired = pf.weightUnbalanced()
        wn = np.array([p.weight for p in pf.particles])
        np.random.seed(100 + n)
        pf.resample()
        # resample() deep-copies our stand-ins; recover the indices from a tag
        np.random.seed(100 + n)
        u = np.random.random_sample(n)
        np.random.seed(100 + n)
        idx = np.random.choice(np.arange(n), n, p=wn)
        out["w0_%d" % n], out["wn_%d" % n], out["fired_%d" % n] = w0, wn, np.array([fired])
        out["u_%d" % n], out["idx_%d" % n] = u, idx
    # degenerate cases for the trigger (FastSlam.py:37)
    for n in (4, 10, 15):
        for slot in (0, n - 1):
            pf = fsmod.ParticleFilter.__new__(fsmod.ParticleFilter)
            pf.numParticles = n
            w = np.full(n, 1e-30)
            w[slot] = 1.0
            pf.particles = [P(float(x)) for x in w]
            out["degenerate_%d_%d" % (n, slot)] = np.array([pf.weightUnbalanced()])
    return out


def main_full():
    """Whole-log goldens (F4 / VERDICT item 1f): poses, confidences, count sums of the deterministic driver."""
    frames = load_frames("intel_gfs", 10 ** 6)
    with open(os.path.join(REF, "DataSet/PreprocessedData", "intel_corrected_log")) as f:
        gt = json.load(f)["map"]
    assert sorted(gt.keys()) == [fr["key"] for fr in frames]
    np.savez_compressed(
        os.path.join(HERE, "intel_full.npz"),
        keys=np.array([float(fr["key"]) for fr in frames]),
        poses=np.array([[fr["x"], fr["y"], fr["theta"]] for fr in frames], dtype=np.float64),
        ranges=np.array([fr["range"] for fr in frames], dtype=np.float64),
        truth=np.array([[gt[fr["key"]]["x"], gt[fr["key"]]["y"], gt[fr["key"]]["theta"]] for fr in frames], dtype=np.float64))
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    out = {}
    for tag, ogArgs, smArgs in (("c3", (72, 72, init, 0.05, np.pi, 180, 10, 0.25), (1.5, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 5)),
                                ("ref02", (72, 72, init, 0.02, np.pi, 180, 10, 5 * 0.02), (1.4, 0.25, 2, 0.1, 0.25, 0.3, 0.15, 5))):
        r = run_deterministic(frames, ogArgs, smArgs, len(frames), set())
        out[tag + "_poses"], out[tag + "_confs"] = r["poses"], r["confs"]
        out[tag + "_sums"] = np.array([float(r["visited"].astype(np.float64).sum()), float(r["total"].astype(np.float64).sum()),
                                       float(len(r["cells"]))])
        out[tag + "_sha"] = np.frombuffer(bytes.fromhex(sha(r["poses"])), dtype=np.uint8)
        print(tag, "trajectory sha", sha(r["poses"]), "touched cells", len(r["cells"]))
    np.savez_compressed(os.path.join(HERE, "det_intel_full.npz"), **out)


def main():
    if "--full" in sys.argv:
        return main_full()
    frames = load_frames()
    with open(os.path.join(HERE, "intel_gfs_head.json"), "w") as f:
        json.dump({"source": "DataSet/PreprocessedData/intel_gfs, first %d readings (sorted keys)" % N_FRAMES,
                   "frames": frames}, f)
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}

    # c3 geometry
    og_c3 = (50, 50, init, 0.05, np.pi, 180, 10, 0.25)
    sm_c3 = (1.5, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 5)
    np.savez_compressed(os.path.join(HERE, "det_c3.npz"), **run_deterministic(frames, og_c3, sm_c3, 30, {2, 9, 17}))

    # reference defaults (Utils/ScanMatcher_OGBased.py:292-294) on a pre-sized 50 m map
    og_02 = (50, 50, init, 0.02, np.pi, 180, 10, 5 * 0.02)
    sm_02 = (1.4, 0.25, 2, 0.1, 0.25, 0.3, 0.15, 5)
    np.savez_compressed(os.path.join(HERE, "det_ref02.npz"), **run_deterministic(frames, og_02, sm_02, 12, {2}))

    # FastSLAM, argument order of Algorithm/FastSlam.py:204
    ogp = [50, 50, init, 0.05, np.pi, 10, 180, 0.25]
    np.savez_compressed(os.path.join(HERE, "pf_c3.npz"), **run_fastslam(frames, ogp, list(sm_c3), 3, 22, seed=0))

    np.savez_compressed(os.path.join(HERE, "update_c3.npz"), **run_update_only(frames, og_c3, 12))

    # MIT CSAIL log: 361 beams (numSpokes 722), poses near x = 576 m (different float noise in the index maps)
    cs = load_frames("csail_gfs", 20)
    with open(os.path.join(HERE, "csail_gfs_head.json"), "w") as f:
        json.dump({"source": "DataSet/PreprocessedData/csail_gfs, first 20 readings (sorted keys)", "frames": cs}, f)
    cinit = {"x": cs[0]["x"], "y": cs[0]["y"]}
    og_cs = (50, 50, cinit, 0.05, np.pi, 361, 10, 0.25)
    np.savez_compressed(os.path.join(HERE, "det_csail.npz"), **run_deterministic(cs, og_cs, sm_c3, 20, {3, 12}))
    np.savez_compressed(os.path.join(HERE, "resample.npz"), **run_resample())
    for fn in sorted(os.listdir(HERE)):
        p = os.path.join(HERE, fn)
        if os.path.isfile(p):
            print("%-24s %8d bytes" % (fn, os.path.getsize(p)))


if __name__ == "__main__":
    main()
