"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest

from conftest import load_golden, reading, dense_counts
from oracle import slam_oracle as O


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


OG_C3 = lambda init: (50, 50, init, 0.05, np.pi, 180, 10, 0.25)
SM_C3 = (1.5, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 5)
OG_02 = lambda init: (50, 50, init, 0.02, np.pi, 180, 10, 5 * 0.02)
SM_02 = (1.4, 0.25, 2, 0.1, 0.25, 0.3, 0.15, 5)


def drive_deterministic(frames, ogArgs, smArgs, n, traceSteps=()):
    """Same loop as Utils/ScanMatcher_OGBased.py:226-256, on the oracle's classes."""
    og = O.OccupancyGrid(*ogArgs)
    sm = O.ScanMatcher(og, *smArgs)
    poses, confs, traces = [], [], {}
    xT, yT = [], []
    for count, fr in enumerate(frames[:n], start=1):
        cur = reading(fr)
        if count == 1:
            prevRawTh = prevMatchedTh = None
            matched, conf = cur, 1
        else:
            ex, ey, eth, dist, estTh, rawTh = O.propose_pose(cur, prevMatched, prevRaw, prevRawTh, prevMatchedTh)
            sm.trace = [] if count in traceSteps else None
            matched, conf = sm.matchScan({'x': ex, 'y': ey, 'theta': eth, 'range': cur['range']}, dist, estTh, count)
            if sm.trace is not None:
                traces[count] = sm.trace
            prevRawTh = rawTh
            prevMatchedTh = O.moving_heading(matched['x'], matched['y'], xT[-1], yT[-1])
        og.updateOccupancyGrid(matched)
        xT.append(matched['x']); yT.append(matched['y'])
        prevMatched, prevRaw = matched, cur
        poses.append([matched['x'], matched['y'], matched['theta']])
        confs.append(conf)
    return og, np.array(poses), np.array(confs, dtype=np.float64), traces


@pytest.mark.parametrize("name,og,sm,n,steps", [("det_c3.npz", OG_C3, SM_C3, 30, (2, 9, 17)),
                                                ("det_ref02.npz", OG_02, SM_02, 12, (2,))])
def test_deterministic_driver_matches_reference(frames, name, og, sm, n, steps):
    g = load_golden(name)
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    grid, poses, confs, traces = drive_deterministic(frames, og(init), sm, n, steps)
    assert np.array_equal(poses, g["poses"])
    assert np.array_equal(confs, g["confs"])
    G = int(g["G"][0])
    v, t = dense_counts(G, g["cells"], g["visited"], g["total"])
    assert np.array_equal(grid.occupancyGridVisited, v) and np.array_equal(grid.occupancyGridTotal, t)
    for c in steps:
        for k, stage in enumerate(("coarse", "fine")):
            tr = traces[c][k]
            tag = "c%d_%s" % (c, stage)
            assert np.array_equal(tr["vol"], g[tag + "_vol"])
            assert np.array_equal(sha(tr["prob"]), g[tag + "_prob_sha"])


def test_fastslam_seeded_matches_reference(frames):
    g = load_golden("pf_c3.npz")
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    np.random.seed(0)
    pf = O.ParticleFilter(3, [50, 50, init, 0.05, np.pi, 10, 180, 0.25], list(SM_C3))
    for count, fr in enumerate(frames[:22], start=1):
        pf.updateParticles(reading(fr), count)
        raw = [p.weight for p in pf.particles]
        fired = pf.weightUnbalanced()
        if fired:
            pf.resample()
        assert fired == bool(g["resampled"][count - 1])
        poses = np.array([[p.prevMatchedReading[k] for k in ("x", "y", "theta")] for p in pf.particles])
        assert np.array_equal(poses, g["poses"][count - 1]), count
        assert np.array_equal(np.array(raw, dtype=np.float64), g["weights"][count - 1][0])
        assert np.array_equal(np.array([p.weight for p in pf.particles]), g["weights"][count - 1][1])
    assert np.random.random_sample() == g["next_uniform"][0]          # RNG stream position
    G = int(g["G"][0])
    for i, p in enumerate(pf.particles):
        v, t = dense_counts(G, g["cells%d" % i], g["visited%d" % i], g["total%d" % i])
        assert np.array_equal(p.og.occupancyGridVisited, v) and np.array_equal(p.og.occupancyGridTotal, t)


def test_update_only_and_tables_match_reference(frames):
    g = load_golden("update_c3.npz")
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    og = O.OccupancyGrid(*OG_C3(init))
    geom = og.geom
    assert np.array_equal(sha(geom.gridX), g["gridX_sha"]) and np.array_equal(sha(geom.gridY), g["gridY_sha"])
    assert [geom.numSpokes, geom.spokesStartIdx] == list(g["numSpokes"][:2])
    sizes = np.bincount(geom.sector.reshape(-1), minlength=geom.numSpokes)
    assert np.array_equal(sizes, g["spoke_sizes"])
    # per-spoke lists in the reference's argwhere (row-major) order
    assert np.array_equal(sha(geom.radius[geom.sector == 7]), g["spoke7_r_sha"])
    lx = np.broadcast_to(geom.localAxis[None, :], geom.sector.shape)
    assert np.array_equal(sha(lx[geom.sector == 200]), g["spoke200_x_sha"])
    for fr in frames[:12]:                      # raw odometry poses: includes off-lattice poses
        og.updateOccupancyGrid(reading(fr))
    v, t = dense_counts(int(g["G"][0]), g["cells"], g["visited"], g["total"])
    assert np.array_equal(og.occupancyGridVisited, v) and np.array_equal(og.occupancyGridTotal, t)


def test_weights_trigger_and_resample_known_answers():
    g = load_golden("resample.npz")
    for n in (5, 10, 15, 64):
        wn = O.normalize([float(w) for w in g["w0_%d" % n]])
        assert np.array_equal(np.array(wn), g["wn_%d" % n])
        assert O.unbalanced(wn)[0] == bool(g["fired_%d" % n][0])
        assert np.array_equal(O.resample_indices(wn, g["u_%d" % n]), g["idx_%d" % n])
    for n in (4, 10, 15):
        for slot in (0, n - 1):
            w = np.full(n, 1e-30); w[slot] = 1.0
            assert O.unbalanced(O.normalize([float(x) for x in w]))[0] == bool(g["degenerate_%d_%d" % (n, slot)][0])


def test_csail_361_beam_driver_matches_reference():
    """MIT CSAIL readings: 361 beams (numSpokes 722), poses ~576 m from the origin (index-map float noise)."""
    import json, os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "csail_gfs_head.json")) as f:
        cs = json.load(f)["frames"]
    g = load_golden("det_csail.npz")
    init = {"x": cs[0]["x"], "y": cs[0]["y"]}
    grid, poses, confs, traces = drive_deterministic(cs, (50, 50, init, 0.05, np.pi, 361, 10, 0.25), SM_C3, 20, (3, 12))
    assert grid.geom.numSpokes == 722 and grid.geom.K == 361
    assert np.array_equal(poses, g["poses"]) and np.array_equal(confs, g["confs"])
    v, t = dense_counts(int(g["G"][0]), g["cells"], g["visited"], g["total"])
    assert np.array_equal(grid.occupancyGridVisited, v) and np.array_equal(grid.occupancyGridTotal, t)
    for c in (3, 12):
        for k, stage in enumerate(("coarse", "fine")):
            assert np.array_equal(traces[c][k]["vol"], g["c%d_%s_vol" % (c, stage)])
            assert np.array_equal(sha(traces[c][k]["prob"]), g["c%d_%s_prob_sha" % (c, stage)])
