"""GPU parity, second set (round-2 review items): diverged large batches against the oracle slot by slot, the CSAIL
361-beam golden, BASELINE config-5 geometry, the global-slot (SLOW) plan in sampling mode, the FastSLAM.step facade,
the whole 910-reading Intel log, and translation invariance.  Tolerances as in test_gpu_parity.py: poses, indices and
map counts exact; score volumes / likelihood fields bit-exact; confidences and weights rtol 1e-12 / 1e-9 (exp)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden, reading, dense_counts
from oracle import slam_oracle as O
from test_gpu_parity import S, drive, OG_C3, SM_C3, OG_02, SM_02, CONF_RTOL, _poses   # noqa: F401  (S is a fixture)

pytestmark = pytest.mark.gpu


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def _oracle_slots(spec, scene, slots, uniforms, steps):
    """Oracle particles for the given slots of a batched run: same pre-warmed map, slot i fed uniforms[step][i]."""
    mapX, mapY, initXY, unit, fov, maxRange, K, wall = spec["og"]
    geom = O.GridGeometry(mapX, mapY, initXY, unit, fov, K, maxRange, wall)
    out = {}
    for i in slots:
        p = O.Particle(spec["og"], spec["sm"], geometry=geom)
        for fr in scene["warm"]:
            p.og.updateOccupancyGrid(fr)
        for count, fr in enumerate(scene["frames"][:steps], start=1):
            p.update(fr, count, uniform=None if count == 1 else float(uniforms[count][i]))
        out[i] = p
    return out


def _warm(S, pf, scene):
    og = S.OccupancyGrid(*pf.geom.args, _geometry=pf.geom)
    for fr in scene["warm"]:
        og.updateOccupancyGrid(fr)
    pf.grids.copy_(og.device_grid.unsqueeze(0).expand_as(pf.grids))


def test_diverged_1024_batch_slots_match_oracle(S):
    """BASELINE config 3 at full size, SAMPLING mode: after 4 steps the 1024 particles have diverged; 10 scattered slots
    (first / last of the grid-stride waves included) are compared with oracle particles fed the same uniforms."""
    from slam_2d_lidar_scan_b200 import synthetic
    spec = synthetic.config("c3")
    n, steps = 1024, 4
    scene = synthetic.make_scene(seed=0, steps=steps + 1, K=180, fov=np.pi, unit=0.05)
    rng = np.random.RandomState(77)
    uniforms = {c: rng.random_sample(n) for c in range(2, steps + 1)}
    pf = S.ParticleFilter(n, spec["og"], spec["sm"])
    _warm(S, pf, scene)
    for count, fr in enumerate(scene["frames"][:steps], start=1):
        pf._update(0, n, fr, count, uniforms=uniforms.get(count))
    rawW = pf.weights.cpu().numpy().copy()
    assert not pf.weightUnbalanced()
    poses, idx = pf.poses(), pf._idx.cpu().numpy()
    assert len(np.unique(poses, axis=0)) >= 8            # really diverged (the coarse softmax is peaked)
    slots = [0, 1, 147, 148, 295, 511, 700, 887, 1022, 1023]
    ref = _oracle_slots(spec, scene, slots, uniforms, steps)
    for i in slots:
        p = ref[i]
        want = [p.prevMatchedReading[k] for k in ("x", "y", "theta")]
        assert poses[i].tolist() == want, i
        assert rawW[i] == pytest.approx(p.weight, rel=1e-9, abs=0)
        g = pf.grids[i].cpu().numpy()
        G = pf.geom.G
        assert np.array_equal(g[:, :G, 0].astype(np.float64), p.og.occupancyGridVisited), i
        assert np.array_equal(g[:, :G, 1].astype(np.float64), p.og.occupancyGridTotal), i
    assert abs(pf.weights.sum().item() - 1.0) < 1e-12
    assert (idx[:, 0] >= 0).all() and (idx[:, 0] < 36).all()


def test_csail_361_beam_driver_matches_reference_golden(S):
    """MIT CSAIL readings: 361 beams (numSpokes 722, 70 rotations, 512-key sorts), poses ~576 m from the origin."""
    with open(os.path.join(GOLDEN, "csail_gfs_head.json")) as f:
        cs = json.load(f)["frames"]
    g = load_golden("det_csail.npz")
    init = {"x": cs[0]["x"], "y": cs[0]["y"]}
    grid, poses, confs, traces = drive(S, cs, (50, 50, init, 0.05, np.pi, 361, 10, 0.25), SM_C3, 20, (3, 12))
    assert grid.geom.numSpokes == 722
    assert np.array_equal(poses, g["poses"])
    np.testing.assert_allclose(confs, g["confs"], rtol=CONF_RTOL, atol=0)
    v, t = dense_counts(int(g["G"][0]), g["cells"], g["visited"], g["total"])
    assert np.array_equal(grid.occupancyGridVisited, v) and np.array_equal(grid.occupancyGridTotal, t)
    for c in (3, 12):
        for stage in ("coarse", "fine"):
            tag = "c%d_%s" % (c, stage)
            assert np.array_equal(traces[c][stage + "_vol"], g[tag + "_vol"]), tag
            assert np.array_equal(sha(traces[c][stage + "_prob"]), g[tag + "_prob_sha"]), tag


def test_c5_geometry_fov_pi_360_beams_matches_oracle(S):
    """BASELINE config 5 geometry: 360 beams over FOV pi (numSpokes 720, 70 rotations), 2001^2 lattice, sampling mode,
    3 particles against the oracle (poses / maps exact, weights 1e-9)."""
    from slam_2d_lidar_scan_b200 import synthetic
    init = {"x": 0.0, "y": 0.0}
    spec = dict(og=[100, 100, init, 0.05, np.pi, 10, 360, 0.25], sm=[1.5, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 5])
    scene = synthetic.make_scene(seed=2, steps=5, K=360, fov=np.pi, unit=0.05)
    n, steps = 3, 4
    rng = np.random.RandomState(5)
    uniforms = {c: rng.random_sample(n) for c in range(2, steps + 1)}
    pf = S.ParticleFilter(n, spec["og"], spec["sm"])
    assert pf.geom.G == 2001 and pf.geom.numSpokes == 720 and pf.engine.volume_shape(0) == (70, 13, 13)
    _warm(S, pf, scene)
    for count, fr in enumerate(scene["frames"][:steps], start=1):
        pf._update(0, n, fr, count, uniforms=uniforms.get(count))
    rawW = pf.weights.cpu().numpy().copy()
    pf.weightUnbalanced()
    ref = _oracle_slots(spec, scene, range(n), uniforms, steps)
    poses = pf.poses()
    for i in range(n):
        p = ref[i]
        assert poses[i].tolist() == [p.prevMatchedReading[k] for k in ("x", "y", "theta")], i
        assert rawW[i] == pytest.approx(p.weight, rel=1e-9, abs=0)
        assert np.array_equal(pf.particles[i].og.occupancyGridTotal, p.og.occupancyGridTotal)
        assert np.array_equal(pf.particles[i].og.occupancyGridVisited, p.og.occupancyGridVisited)


def test_c5_workload_full_circle_2001_lattice_matches_oracle(S):
    """The bench's c5 workload (360 beams over 2*pi, 2001^2 lattice, 0.05 m cells), sampling mode, 2 particles."""
    from slam_2d_lidar_scan_b200 import synthetic
    spec = synthetic.config("c5")
    scene = synthetic.make_scene(seed=0, steps=4, K=360, fov=2 * np.pi, unit=0.05)
    n, steps = 2, 3
    rng = np.random.RandomState(6)
    uniforms = {c: rng.random_sample(n) for c in range(2, steps + 1)}
    pf = S.ParticleFilter(n, spec["og"], spec["sm"])
    _warm(S, pf, scene)
    for count, fr in enumerate(scene["frames"][:steps], start=1):
        pf._update(0, n, fr, count, uniforms=uniforms.get(count))
    pf.weightUnbalanced()
    ref = _oracle_slots(spec, scene, range(n), uniforms, steps)
    for i in range(n):
        p = ref[i]
        assert pf.poses()[i].tolist() == [p.prevMatchedReading[k] for k in ("x", "y", "theta")], i
        assert np.array_equal(pf.particles[i].og.occupancyGridTotal, p.og.occupancyGridTotal)


def test_global_slot_plan_in_sampling_mode_matches_oracle(S, frames):
    """Reference default geometry (0.02 m cells, 1241^2 fine field -> the SLOW global-slot plan) with matchMax=False:
    both sides draw from numpy's global RandomState, so they run one after the other from the same seed."""
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}

    def run(mod, steps=8):
        np.random.seed(13)
        og = mod.OccupancyGrid(*OG_02(init))
        sm = mod.ScanMatcher(og, *SM_02)
        poses, confs = [], []
        xT, yT = [], []
        for count, fr in enumerate(frames[:steps], start=1):
            cur = reading(fr)
            if count == 1:
                prevRawTh = prevMatchedTh = None
                matched, conf = cur, 1
            else:
                ex, ey, eth, dist, estTh, rawTh = O.propose_pose(cur, prevMatched, prevRaw, prevRawTh, prevMatchedTh)
                matched, conf = sm.matchScan({'x': ex, 'y': ey, 'theta': eth, 'range': cur['range']}, dist, estTh, count,
                                             matchMax=False)
                prevRawTh = rawTh
                prevMatchedTh = O.moving_heading(matched['x'], matched['y'], xT[-1], yT[-1])
            og.updateOccupancyGrid(matched)
            xT.append(matched['x']); yT.append(matched['y'])
            prevMatched, prevRaw = matched, cur
            poses.append([matched['x'], matched['y'], matched['theta']])
            confs.append(conf)
        return og, np.array(poses), np.array(confs, dtype=np.float64), np.random.random_sample()
    og, poses, confs, nxt = run(S)
    plan = (S._native.C.c_int * 8)()
    S._native.lib.slam_matcher_plan(S.ScanMatcher(og, *SM_02).engine.handle, 1, S._native.C.byref(plan))
    assert plan[2] == 0                                   # bitmaps in the global slot: this IS the SLOW plan
    rog, rposes, rconfs, rnxt = run(O)
    assert np.array_equal(poses, rposes) and nxt == rnxt
    np.testing.assert_allclose(confs, rconfs, rtol=CONF_RTOL, atol=0)
    assert np.array_equal(og.occupancyGridTotal, rog.occupancyGridTotal)
    assert len(np.unique(poses[:, 2] - np.array([f["theta"] for f in frames[:8]]))) > 1


def test_fastslam_step_facade_matches_oracle_loop(S, frames):
    """FastSLAM.step(reading) == the loop body of Algorithm/FastSlam.py:159-162 (update, trigger, resample)."""
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    ogp, smp = [30, 30, init, 0.1, np.pi, 10, 180, 0.5], [1.0, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 2]
    np.random.seed(8)
    fs = S.FastSLAM(5, ogp, smp)
    fired = [fs.step(reading(fr)) for fr in frames[:10]]
    got = fs.pf.poses()
    nxt = np.random.random_sample()
    np.random.seed(8)
    ref = O.ParticleFilter(5, ogp, smp)
    want = []
    for count, fr in enumerate(frames[:10], start=1):
        ref.updateParticles(reading(fr), count)
        f = ref.weightUnbalanced()
        if f:
            ref.resample()
        want.append(f)
    assert fired == want and fs.count == 10 and fs.resampled == want
    assert np.array_equal(got, _poses(ref)) and nxt == np.random.random_sample()
    np.testing.assert_allclose(fs.pf.weights.cpu().numpy(), [p.weight for p in ref.particles], rtol=1e-9, atol=0)
    for i in (0, 4):
        assert np.array_equal(fs.pf.particles[i].og.occupancyGridTotal, ref.particles[i].og.occupancyGridTotal)


@pytest.mark.parametrize("tag,og,sm", [("c3", lambda i: (72, 72, i, 0.05, np.pi, 180, 10, 0.25), SM_C3),
                                       ("ref02", lambda i: (72, 72, i, 0.02, np.pi, 180, 10, 5 * 0.02), SM_02)])
def test_whole_intel_log_matches_reference_golden(S, tag, og, sm):
    """All 910 readings of intel_gfs through the deterministic driver on a pre-sized 72 m map: every pose, confidence and
    the final count sums equal the unmodified reference's (tests/golden/make_golden.py --full)."""
    full, g = load_golden("intel_full.npz"), load_golden("det_intel_full.npz")
    fr = [dict(x=float(p[0]), y=float(p[1]), theta=float(p[2]), range=r.tolist()) for p, r in zip(full["poses"], full["ranges"])]
    init = {"x": fr[0]["x"], "y": fr[0]["y"]}
    grid, poses, confs, _ = drive(S, fr, og(init), sm, len(fr))
    assert np.array_equal(poses, g[tag + "_poses"])
    assert np.array_equal(sha(poses), g[tag + "_sha"])
    np.testing.assert_allclose(confs, g[tag + "_confs"], rtol=CONF_RTOL, atol=0)
    dg = grid.device_grid[:, :grid.geom.G].to(torch.float64)
    mask = (dg[..., 0] != 1) | (dg[..., 1] != 2)          # the golden sums run over the touched cells
    got = [float(dg[..., 0][mask].sum().item()), float(dg[..., 1][mask].sum().item()), float(mask.sum().item())]
    assert got == g[tag + "_sums"].tolist()


def test_translation_by_whole_cells_shifts_the_result(S, frames):
    """SURVEY section 4 property: moving the map origin and every pose by k*unit moves the matched pose by k*unit and
    leaves the argmax indices and the map counts unchanged (the binary fraction 2^-4 keeps the shift exact)."""
    outs = []
    for shift in (0.0, 64 * 0.0625):
        init = {"x": 1.0 + shift, "y": -2.0 + shift}
        og = S.OccupancyGrid(40, 40, init, 0.0625, np.pi, 180, 10, 0.25)
        sm = S.ScanMatcher(og, 1.5, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 5)
        idxs, poses = [], []
        for count, fr in enumerate(frames[:6], start=1):
            rd = reading(fr)
            rd["x"], rd["y"] = init["x"] + 0.0625 * round((fr["x"] - frames[0]["x"]) / 0.0625), \
                init["y"] + 0.0625 * round((fr["y"] - frames[0]["y"]) / 0.0625)
            m, c = sm.matchScan(rd, 0.1, None, count)
            if count > 1:
                idxs.append(sm.lastIdx)
            og.updateOccupancyGrid(m)
            poses.append([m["x"] - shift, m["y"] - shift, m["theta"]])
        outs.append((idxs, np.array(poses), og.occupancyGridTotal.copy()))
    assert outs[0][0] == outs[1][0]
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])


@pytest.mark.gpu
def test_sharding_is_invisible_across_two_gpus():
    """Two ranks (NCCL) vs the unsharded filter, bit for bit, incl. a forced cross-rank resample.
    Needs two GPUs on the box; skipped otherwise."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29400 + os.getpid() % 500
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(root, "tools", "check_sharded.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def _oracle_stage(ref, mp, x, y, th, rng, ul, sigma, miss, searchRadius, estDist, estTheta, fine, matchMax, prob=None):
    """One stage of the oracle, piece by piece -> (xr0, yr0, probSP, convTotal, (it, iy, ix), conf)."""
    g = ref.geom
    R = 1.1 * g.maxRange + mp[0]
    bx, by, space = O.occupancy_scatter(g, ref.occupancyGridVisited, ref.occupancyGridTotal, x, y, ul, R, miss)
    if prob is None:
        prob = O.likelihood_field(space, sigma)
    px, py = O.beam_endpoints(g, x, y, th, rng)
    axis = O.offset_axis(searchRadius, ul)
    if fine:
        rv = tw = np.zeros((axis.shape[0], axis.shape[0]))
    else:
        rv, tw = O.motion_priors(axis, ul, estDist, estTheta, mp[3], mp[4], mp[5])
    thetas = O.theta_offsets(g, mp[1])
    with np.errstate(invalid="ignore"):
        vol = O.correlation_volume(prob, px, py, x, y, bx, by, ul, thetas, axis, rv, tw)
        idx, conf = O.choose_pose(vol, matchMax)
    pose = (x + axis[idx[2]] * ul, y + axis[idx[1]] * ul, th + thetas[idx[0]])
    return bx, by, prob, vol, pose, conf


@pytest.mark.gpu
def test_grouped_map_update_many_headings_and_a_patch_leaving_the_map(S):
    """slam_update_grid groups the particles by sector shift: 70 particles with 40 distinct headings (several work items,
    partial groups), one pose half a cell off the lattice (general path in the same launch) and one patch that leaves
    the lattice (cells outside are skipped, SLAM_ST_SCAN_OUTSIDE_MAP is raised for that particle only); counts
    identical to the oracle's per-particle OccupancyGrid.updateOccupancyGrid (OccupancyGrid.py:127-152)."""
    from slam_2d_lidar_scan_b200.engine import update_grids
    nat = S._native
    init = {"x": 0.0, "y": 0.0}
    args = (30, 30, init, 0.1, np.pi, 180, 10, 0.5)
    og = S.OccupancyGrid(*args)
    geom = og.geom
    rng = np.random.default_rng(7)
    n = 70
    ranges = np.round(rng.uniform(1.0, 12.0, 180), 2)
    poses = np.zeros((n, 3))
    poses[:, 0] = np.round(rng.uniform(-3, 3, n), 1)            # on the lattice: pure shifts
    poses[:, 1] = np.round(rng.uniform(-3, 3, n), 1)
    poses[:, 2] = (np.arange(n) % 40) * (2 * np.pi / 360) * 1.7 - 0.6
    poses[5, 0] += 0.05                                         # half a cell off: rint ties -> general path
    poses[9, :2] = (9.0, -2.0)                                  # patch reaches x = 19 > 15: leaves the lattice
    dev = geom.device
    grids = geom.new_grids(n)
    status = torch.zeros(n, dtype=torch.int32, device=dev)
    d_r = torch.from_numpy(ranges).to(dev)
    d_p = torch.from_numpy(poses).to(dev)
    for _ in range(2):
        update_grids(geom, grids, n, d_r, d_p, status)
    torch.cuda.synchronize()
    st = status.cpu().numpy()
    assert st[9] == nat.ST_SCAN_OUTSIDE_MAP and not st[np.arange(n) != 9].any()
    G = geom.G
    host = grids.cpu().numpy()[:, :, :G, :]
    for i in (0, 3, 5, 17, 39, 40, 41, 69):
        ref = O.OccupancyGrid(*args)
        rd = {"x": poses[i, 0], "y": poses[i, 1], "theta": poses[i, 2], "range": list(ranges)}
        ref.updateOccupancyGrid(rd)
        ref.updateOccupancyGrid(rd)
        assert np.array_equal(host[i, :, :, 0], ref.occupancyGridVisited), i       # both [row = y][col = x]
        assert np.array_equal(host[i, :, :, 1], ref.occupancyGridTotal), i


def test_stage_methods_match_the_oracle_stage_by_stage(S, frames):
    """The reference's public stage API (ScanMatcher_OGBased.py:20-45, 81-176): frameSearchSpace ->
    searchToMatch (coarse, argmax and sampled, with and without heading prior) -> frameSearchSpace -> searchToMatch
    (fine), each through the stage-level C entries, against the oracle's stages on the same map: probSP and convTotal
    bit for bit, same pose, same position in numpy's RNG stream."""
    from scipy.ndimage import gaussian_filter
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    ogp = (30, 30, init, 0.1, np.pi, 180, 10, 0.5)
    smp = (1.0, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 2)
    og, ref = S.OccupancyGrid(*ogp), O.OccupancyGrid(*ogp)
    sm = S.ScanMatcher(og, *smp)
    for fr in frames[:4]:
        og.updateOccupancyGrid(reading(fr)); ref.updateOccupancyGrid(reading(fr))
    fr = reading(frames[4])
    x, y, th, rng = fr["x"], fr["y"], fr["theta"], np.asarray(fr["range"], dtype=np.float64)
    cstep, csig = smp[7] * og.unitGridSize, smp[2] / smp[7]
    xr, yr, prob = sm.frameSearchSpace(x, y, cstep, csig, smp[6])
    bx, by, prob2, _, _, _ = _oracle_stage(ref, smp, x, y, th, rng, cstep, csig, smp[6], smp[0], 0.12, None, False, True)
    assert xr[0] == bx and yr[0] == by and np.array_equal(prob, prob2)
    pose = None
    for matchMax, mt in ((True, None), (False, 0.4), (False, None)):
        np.random.seed(11)
        got = sm.searchToMatch(prob, x, y, th, rng, xr, yr, smp[0], smp[1], cstep, 0.12, mt, fineSearch=False, matchMax=matchMax)
        after = np.random.random_sample()
        np.random.seed(11)
        _, _, _, vol, pose, conf = _oracle_stage(ref, smp, x, y, th, rng, cstep, csig, smp[6], smp[0], 0.12, mt, False, matchMax)
        assert after == np.random.random_sample()                               # same number of draws consumed
        assert np.array_equal(got[3], vol)                                       # convTotal
        assert (got[2]["x"], got[2]["y"], got[2]["theta"]) == pose
        assert got[4] == pytest.approx(conf, rel=CONF_RTOL)
        assert len(got[0]) == len(got[1]) == int((rng < 10).sum())
    fmiss = smp[6] ** (2 / smp[7])
    xr, yr, prob = sm.frameSearchSpace(pose[0], pose[1], og.unitGridSize, smp[2], fmiss)
    _, _, prob2, vol, fpose, _ = _oracle_stage(ref, smp, pose[0], pose[1], pose[2], rng, og.unitGridSize, smp[2], fmiss, cstep,
                                               0.12, None, True, True)
    assert np.array_equal(prob, prob2)
    got = sm.searchToMatch(prob, pose[0], pose[1], pose[2], rng, xr, yr, cstep, smp[1], og.unitGridSize, 0.12, None, fineSearch=True)
    assert np.array_equal(got[3], vol) and (got[2]["x"], got[2]["y"], got[2]["theta"]) == fpose
    # generateProbSearchSpace on an arbitrary array == scipy + min + clamp; helpers == the reference expressions
    a = np.random.default_rng(2).normal(-1.0, 0.7, (57, 83))
    want = gaussian_filter(a, sigma=1.3)
    want[want > 0.5 * want.min()] = 0
    assert np.array_equal(sm.generateProbSearchSpace(a, 1.3), want)
    px, py = sm.covertMeasureToXY(x, y, th, rng)
    px2, py2 = O.beam_endpoints(ref.geom, x, y, th, rng)
    assert np.array_equal(px, px2) and np.array_equal(py, py2)
    qx, qy = sm.rotate((x, y), (px, py), 0.1)
    assert np.array_equal(qx, x + np.cos(0.1) * (px - x) - np.sin(0.1) * (py - y))
    xi, yi = sm.convertXYToSearchSpaceIdx(px, py, xr[0], yr[0], 0.1)
    assert np.array_equal(xi, ((px - xr[0]) / 0.1).astype(int)) and np.array_equal(yi, ((py - yr[0]) / 0.1).astype(int))


@pytest.mark.gpu
def test_nan_score_volume_argmax_returns_numpys_first_nan(S, frames):
    """numpy's argmax returns the first NaN; np.random.choice raises on NaN probabilities (:134-138)."""
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    ogp = (30, 30, init, 0.1, np.pi, 180, 10, 0.5)
    smp = (1.0, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 2)
    og, ref = S.OccupancyGrid(*ogp), O.OccupancyGrid(*ogp)
    sm = S.ScanMatcher(og, *smp)
    fr = reading(frames[1])
    x, y, th, rng = fr["x"], fr["y"], fr["theta"], np.asarray(fr["range"], dtype=np.float64)
    cstep, csig = smp[7] * og.unitGridSize, smp[2] / smp[7]
    xr, yr, prob = sm.frameSearchSpace(x, y, cstep, csig, smp[6])
    prob = prob.copy()
    prob[prob.shape[0] // 2 - 3: prob.shape[0] // 2 + 3, :] = np.nan
    _, _, _, vol, pose, _ = _oracle_stage(ref, smp, x, y, th, rng, cstep, csig, smp[6], smp[0], 0.1, None, False, True, prob=prob)
    got = sm.searchToMatch(prob, x, y, th, rng, xr, yr, smp[0], smp[1], cstep, 0.1, None, fineSearch=False, matchMax=True)
    assert np.isnan(vol).any() and np.array_equal(np.isnan(got[3]), np.isnan(vol))
    assert (got[2]["x"], got[2]["y"], got[2]["theta"]) == pose
    with pytest.raises(ValueError):
        sm.searchToMatch(prob, x, y, th, rng, xr, yr, smp[0], smp[1], cstep, 0.1, None, fineSearch=False, matchMax=False)


@pytest.mark.gpu
def test_map_expansion_equals_the_presized_lattice(S, frames):
    """F1 (OccupancyGrid.py:59-125): a map that starts at the reference's default 10 m grows (doubling around its
    centre) into exactly the lattice of a map pre-sized to the final length, so the driver loop gives the same poses
    and the same counts as on the pre-sized map -- for the standalone classes and for a particle filter."""
    from slam_2d_lidar_scan_b200.drivers import run_scanmatch
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    smp = (1.4, 0.25, 2, 0.1, 0.25, 0.3, 0.15, 5)
    data = {"%03d" % i: reading(fr) for i, fr in enumerate(frames[:8])}

    def run(length):
        og = S.OccupancyGrid(length, length, init, 0.05, np.pi, 180, 10, 0.25)
        sm = S.ScanMatcher(og, *smp)
        poses, confs = run_scanmatch(data, og, sm)
        return og, poses, confs
    small, p1, c1 = run(10)
    big, p2, c2 = run(small.geom.args[0])
    assert small.geom.args[0] >= 40 and small.geom.G == big.geom.G and small.mapXLim == big.mapXLim
    assert np.array_equal(small.geom.gridX, big.geom.gridX)
    assert np.array_equal(p1, p2) and np.array_equal(c1, c2)
    assert np.array_equal(small.occupancyGridVisited, big.occupancyGridVisited)
    assert np.array_equal(small.occupancyGridTotal, big.occupancyGridTotal)
    # counts written before an expansion are re-homed exactly
    og = S.OccupancyGrid(30, 30, init, 0.1, np.pi, 180, 10, 0.5)
    og.updateOccupancyGrid(reading(frames[0]))
    before, lim = og.occupancyGridTotal, list(og.mapXLim)
    og.checkAndExapndOG(np.array([init["x"] + 20.0]), np.array([init["y"]]))
    off = (og.geom.G - before.shape[0]) // 2
    assert og.mapXLim[1] >= init["x"] + 20.0 and og.mapXLim != lim
    after = og.occupancyGridTotal
    assert np.array_equal(after[off:off + before.shape[0], off:off + before.shape[1]], before)
    assert after.sum() - 2.0 * after.size == before.sum() - 2.0 * before.size
    # particle filter: all lattices grow together
    def runpf(length):
        np.random.seed(9)
        pf = S.ParticleFilter(4, [length, length, init, 0.05, np.pi, 10, 180, 0.25], list(smp))
        for count, fr in enumerate(frames[:5], start=1):
            pf.updateParticles(reading(fr), count)
            pf.weightUnbalanced()
        return pf
    a = runpf(10)
    b = runpf(a.geom.args[0])
    assert a.expansions >= 2 and b.expansions == 0
    assert np.array_equal(a.poses(), b.poses()) and torch.equal(a.weights, b.weights)
    for i in range(4):
        assert np.array_equal(a.particles[i].og.occupancyGridVisited, b.particles[i].og.occupancyGridVisited)
