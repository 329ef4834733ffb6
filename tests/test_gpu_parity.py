"""GPU parity: the CUDA path (through the C ABI) against the reference's golden outputs and the CPU oracle.

Tolerances: pose / argmax / sample / resample indices and map counts are exact (integers, or float64 values
computed from exact indices); the score volume and probSP are bit-exact; confidences and weights go through
exp(), whose device implementation differs from numpy's by <= 1 ulp per term -> rtol 1e-12 on sums.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, reading, dense_counts
from oracle import slam_oracle as O

pytestmark = pytest.mark.gpu

OG_C3 = lambda init: (50, 50, init, 0.05, np.pi, 180, 10, 0.25)
SM_C3 = (1.5, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 5)
OG_02 = lambda init: (50, 50, init, 0.02, np.pi, 180, 10, 5 * 0.02)
SM_02 = (1.4, 0.25, 2, 0.1, 0.25, 0.3, 0.15, 5)
CONF_RTOL = 1e-12


@pytest.fixture(scope="module")
def S():
    import __graft_entry__ as g
    g.build()
    import slam_2d_lidar_scan_b200 as S
    return S


def drive(S, frames, ogArgs, smArgs, n, traceSteps=()):
    """The loop of Utils/ScanMatcher_OGBased.py:226-256 on the CUDA classes."""
    og = S.OccupancyGrid(*ogArgs)
    sm = S.ScanMatcher(og, *smArgs)
    poses, confs, traces, idxs = [], [], {}, []
    xT, yT = [], []
    for count, fr in enumerate(frames[:n], start=1):
        cur = reading(fr)
        if count == 1:
            prevRawTh = prevMatchedTh = None
            matched, conf = cur, 1
        else:
            est, dist, estTh, rawTh = S.updateEstimatedPose(cur, prevMatched, prevRaw, prevRawTh, prevMatchedTh)
            sm.debug = count in traceSteps
            matched, conf = sm.matchScan(est, dist, estTh, count)
            if sm.debug:
                traces[count] = sm.last
            idxs.append(sm.lastIdx)
            prevRawTh = rawTh
            prevMatchedTh = S.getMovingTheta(matched, xT, yT)
        og.updateOccupancyGrid(matched)
        S.updateTrajectory(matched, xT, yT)
        prevMatched, prevRaw = matched, cur
        poses.append([matched['x'], matched['y'], matched['theta']])
        confs.append(conf)
    return og, np.array(poses), np.array(confs, dtype=np.float64), traces


@pytest.mark.parametrize("name,og,sm,n,steps", [("det_c3.npz", OG_C3, SM_C3, 30, (2, 9, 17)),
                                                ("det_ref02.npz", OG_02, SM_02, 12, (2,))])
def test_deterministic_driver_matches_reference_golden(S, frames, name, og, sm, n, steps):
    g = load_golden(name)
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    grid, poses, confs, traces = drive(S, frames, og(init), sm, n, steps)
    assert np.array_equal(poses, g["poses"])
    np.testing.assert_allclose(confs, g["confs"], rtol=CONF_RTOL, atol=0)
    v, t = dense_counts(int(g["G"][0]), g["cells"], g["visited"], g["total"])
    assert np.array_equal(grid.occupancyGridVisited, v) and np.array_equal(grid.occupancyGridTotal, t)
    for c in steps:
        for stage in ("coarse", "fine"):
            tag = "c%d_%s" % (c, stage)
            assert np.array_equal(traces[c][stage + "_vol"], g[tag + "_vol"]), tag
            prob = traces[c][stage + "_prob"]
            assert tuple(prob.shape) == tuple(g[tag + "_prob_shape"])
            stats = g[tag + "_prob_sum"]
            assert prob.min() == stats[1] and float((prob == 0).sum()) == stats[2] and prob.sum() == stats[0]


def test_likelihood_field_bit_exact_vs_oracle(S, frames):
    """probSP of both stages, bit for bit, on a map built from 8 real scans (c3 geometry)."""
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    og = S.OccupancyGrid(*OG_C3(init))
    ref = O.OccupancyGrid(*OG_C3(init))
    for fr in frames[:8]:
        rd = reading(fr)
        rd["x"], rd["y"] = init["x"] + 0.05 * round((rd["x"] - init["x"]) / 0.05), init["y"] + 0.05 * round((rd["y"] - init["y"]) / 0.05)
        og.updateOccupancyGrid(rd)
        ref.updateOccupancyGrid(rd)
    assert np.array_equal(og.occupancyGridVisited, ref.occupancyGridVisited)
    sm, rsm = S.ScanMatcher(og, *SM_C3), O.ScanMatcher(ref, *SM_C3)
    sm.debug, rsm.trace = True, []
    est = reading(frames[8])
    est["x"], est["y"] = init["x"] + 0.15, init["y"] - 0.1
    m, c = sm.matchScan(est, 0.12, 0.4, 2)
    rm, rc = rsm.matchScan(est, 0.12, 0.4, 2)
    for k, stage in enumerate(("coarse", "fine")):
        assert np.array_equal(sm.last[stage + "_prob"], rsm.trace[k]["prob"]), stage
        assert np.array_equal(sm.last[stage + "_vol"], rsm.trace[k]["vol"]), stage
    assert (m["x"], m["y"], m["theta"]) == (rm["x"], rm["y"], rm["theta"])
    assert c == pytest.approx(rc, rel=CONF_RTOL, abs=0)


def test_fastslam_seeded_matches_reference_golden(S, frames):
    g = load_golden("pf_c3.npz")
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    np.random.seed(0)
    pf = S.ParticleFilter(3, [50, 50, init, 0.05, np.pi, 10, 180, 0.25], list(SM_C3))
    for count, fr in enumerate(frames[:22], start=1):
        pf.updateParticles(reading(fr), count)
        raw = pf.weights.cpu().numpy().copy()
        fired = pf.weightUnbalanced()
        if fired:
            pf.resample()
        assert fired == bool(g["resampled"][count - 1])
        assert np.array_equal(pf.poses(), g["poses"][count - 1]), count
        np.testing.assert_allclose(raw, g["weights"][count - 1][0], rtol=1e-9, atol=0)
        np.testing.assert_allclose(pf.weights.cpu().numpy(), g["weights"][count - 1][1], rtol=1e-9, atol=0)
    assert np.random.random_sample() == g["next_uniform"][0]
    G = int(g["G"][0])
    for i, p in enumerate(pf.particles):
        v, t = dense_counts(G, g["cells%d" % i], g["visited%d" % i], g["total%d" % i])
        assert np.array_equal(p.og.occupancyGridVisited, v) and np.array_equal(p.og.occupancyGridTotal, t)
        assert len(p.xTrajectory) == 22


def test_update_only_matches_reference_golden(S, frames):
    """Raw-odometry poses are off-lattice: exercises the general (map-cell-owned) update kernel too."""
    g = load_golden("update_c3.npz")
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    og = S.OccupancyGrid(*OG_C3(init))
    for fr in frames[:12]:
        og.updateOccupancyGrid(reading(fr))
    v, t = dense_counts(int(g["G"][0]), g["cells"], g["visited"], g["total"])
    assert np.array_equal(og.occupancyGridVisited, v) and np.array_equal(og.occupancyGridTotal, t)


def test_update_half_cell_ties_match_oracle(S, frames):
    """Pose exactly half a cell off the lattice: rint's half-to-even collapses neighbouring local cells and numpy's
    fancy += applies once per statement."""
    init = {"x": 0.0, "y": 0.0}
    args = (20, 20, init, 0.1, np.pi, 180, 6, 0.3)
    og, ref = S.OccupancyGrid(*args), O.OccupancyGrid(*args)
    rng = np.random.default_rng(0)
    for k in range(4):
        rd = {"x": 0.05 + 0.1 * k, "y": -0.25, "theta": 0.3 * k, "range": list(np.round(rng.uniform(1, 7, 180), 2))}
        og.updateOccupancyGrid(rd)
        ref.updateOccupancyGrid(rd)
    assert np.array_equal(og.occupancyGridVisited, ref.occupancyGridVisited)
    assert np.array_equal(og.occupancyGridTotal, ref.occupancyGridTotal)
    assert ref.occupancyGridTotal.max() > 2 + 4      # ties really happened somewhere (+3/+4 per scan)


def test_weights_trigger_and_resample_known_answers(S):
    g = load_golden("resample.npz")
    nat = S._native
    dev = torch.device("cuda")
    st = torch.cuda.current_stream().cuda_stream
    for n in (5, 10, 15, 64):
        w = torch.from_numpy(g["w0_%d" % n].copy()).to(dev)
        out = torch.zeros(4, dtype=torch.float64, device=dev)
        nat.check(nat.lib.slam_normalize_weights(n, w.data_ptr(), out.data_ptr(), st))
        assert np.array_equal(w.cpu().numpy(), g["wn_%d" % n])
        assert bool(out[1].item()) == bool(g["fired_%d" % n][0])
        # the fused end-of-step trigger: same normalisation out of place (raw weights untouched) + OR of the status words
        raw = torch.from_numpy(g["w0_%d" % n].copy()).to(dev)
        wn = torch.zeros(n, dtype=torch.float64, device=dev)
        res = torch.zeros(4, dtype=torch.float64, device=dev)
        status = torch.zeros(n, dtype=torch.int32, device=dev)
        status[n // 2], status[n - 1] = 8, 16
        nat.check(nat.lib.slam_step_trigger(n, raw.data_ptr(), wn.data_ptr(), status.data_ptr(), n, res.data_ptr(), st))
        assert np.array_equal(wn.cpu().numpy(), g["wn_%d" % n]) and np.array_equal(raw.cpu().numpy(), g["w0_%d" % n])
        assert res[0].item() == out[0].item() and res[1].item() == out[1].item()
        assert int(res.view(torch.int32)[4].item()) == 24
        u = torch.from_numpy(g["u_%d" % n].copy()).to(dev)
        cdf = torch.zeros(n, dtype=torch.float64, device=dev)
        idx = torch.zeros(n, dtype=torch.int32, device=dev)
        nat.check(nat.lib.slam_resample_indices(n, w.data_ptr(), u.data_ptr(), cdf.data_ptr(), idx.data_ptr(), st))
        assert np.array_equal(idx.cpu().numpy(), g["idx_%d" % n])
    for n in (4, 10, 15):
        for slot in (0, n - 1):
            wv = np.full(n, 1e-30); wv[slot] = 1.0
            w = torch.from_numpy(wv).to(dev)
            out = torch.zeros(4, dtype=torch.float64, device=dev)
            nat.check(nat.lib.slam_normalize_weights(n, w.data_ptr(), out.data_ptr(), st))
            assert bool(out[1].item()) == bool(g["degenerate_%d_%d" % (n, slot)][0])


def _poses(ref):
    return np.array([[p.prevMatchedReading[k] for k in "x y theta".split()] for p in ref.particles])


def test_resample_copies_particles_like_the_oracle(S, frames):
    """Both filters draw from numpy's one global RandomState, so they run one after the other from the same seed."""
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    ogp, smp = [30, 30, init, 0.1, np.pi, 10, 180, 0.5], [1.0, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 2]

    def run(cls):
        np.random.seed(5)
        pf = cls(6, ogp, smp)
        log = []
        for count, fr in enumerate(frames[:9], start=1):
            pf.updateParticles(reading(fr), count)
            log.append(pf.weightUnbalanced())
            if count == 6:
                log.append(pf.poses() if hasattr(pf, "poses") else _poses(pf))
                pf.resample()
                if hasattr(pf, "lastResampleCopies"):       # copy elision: only the extra duplicates move (F2)
                    assert pf.lastResampleCopies == 6 - len(set(int(v) for v in pf.lastResampleIdx))
                log.append(np.array(pf.lastResampleIdx))
                log.append([p.weight for p in pf.particles])
                log.append([p.og.occupancyGridTotal.copy() for p in pf.particles])
            if count == 8:
                pf.resample()
                log.append(np.array(pf.lastResampleIdx))
        log.append([p.og.occupancyGridVisited.copy() for p in pf.particles])
        log.append(pf.poses() if hasattr(pf, "poses") else _poses(pf))
        log.append(np.random.random_sample())
        return log
    got, want = run(S.ParticleFilter), run(O.ParticleFilter)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        if isinstance(a, list):
            assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
        else:
            assert np.array_equal(a, b)


def test_c2_geometry_batch_matches_oracle(S, frames):
    """BASELINE config 2 geometry (unit 0.1, coarse factor 2 -> blur radii 4/8, 11x11 / 5x5 offsets), 16 particles."""
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    ogp, smp = [50, 50, init, 0.1, np.pi, 10, 180, 0.5], [1.1, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 2]
    np.random.seed(11)
    pf = S.ParticleFilter(16, ogp, smp)
    got = []
    for count, fr in enumerate(frames[:14], start=1):
        pf.updateParticles(reading(fr), count)
        pf.weightUnbalanced()
        got.append(pf.poses())
    np.random.seed(11)
    ref = O.ParticleFilter(16, ogp, smp)
    for count, fr in enumerate(frames[:14], start=1):
        ref.updateParticles(reading(fr), count)
        ref.weightUnbalanced()
        assert np.array_equal(got[count - 1], _poses(ref)), count
    np.testing.assert_allclose(pf.weights.cpu().numpy(), [p.weight for p in ref.particles], rtol=1e-9, atol=0)
    for i in (0, 7, 15):
        assert np.array_equal(pf.particles[i].og.occupancyGridTotal, ref.particles[i].og.occupancyGridTotal)
    assert len(np.unique(got[-1][:, 0])) > 1


def test_exact_cdf_walk_equals_certified_parallel_inversion(S, frames):
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    ogp, smp = [50, 50, init, 0.05, np.pi, 10, 180, 0.25], list(SM_C3)
    out = []
    for force in (0, 1):
        np.random.seed(2)
        pf = S.ParticleFilter(32, ogp, smp)
        S._native.lib.slam_matcher_set_debug(pf.engine.handle, None, force)
        for count, fr in enumerate(frames[:16], start=1):
            pf.updateParticles(reading(fr), count)
        out.append((pf.poses(), pf._idx.cpu().numpy(), pf.weights.cpu().numpy()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2], out[1][2])
    assert len(np.unique(out[0][0][:, 2])) > 1          # particles really diverged through sampling


def test_window_outside_map_grows_the_map_or_raises_when_fixed(S, frames):
    """A 20 m map cannot hold the +-12 m search window: the standalone classes grow it like the reference
    (ScanMatcher_OGBased.py:27); a filter told to keep its maps fixed reports the window like numpy would
    (IndexError)."""
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    og = S.OccupancyGrid(20, 20, init, 0.1, np.pi, 180, 10, 0.5)
    sm = S.ScanMatcher(og, 1.0, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 2)
    og.updateOccupancyGrid(reading(frames[0]))
    matched, conf = sm.matchScan(reading(frames[1]), 0.0, None, 2)
    assert og.geom.args[0] == 40 and og.mapXLim[1] - og.mapXLim[0] == pytest.approx(40.0)
    big = S.OccupancyGrid(40, 40, init, 0.1, np.pi, 180, 10, 0.5)
    smBig = S.ScanMatcher(big, 1.0, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 2)
    big.updateOccupancyGrid(reading(frames[0]))
    m2, c2 = smBig.matchScan(reading(frames[1]), 0.0, None, 2)
    assert (matched["x"], matched["y"], matched["theta"], conf) == (m2["x"], m2["y"], m2["theta"], c2)
    pf = S.ParticleFilter(2, [20, 20, init, 0.1, np.pi, 10, 180, 0.5], [1.0, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 2])
    pf.expandMaps = False
    pf.updateParticles(reading(frames[0]), 1)
    pf.weightUnbalanced()
    pf.updateParticles(reading(frames[1]), 2)
    with pytest.raises(IndexError):
        pf.weightUnbalanced()


def test_360_beam_full_circle_scan_matches_oracle(S):
    """BASELINE config 5 flavour: 360 beams over 2*pi (numSpokes 360, spokesStartIdx 270, 512-key sorts), synthetic
    scene, 3 particles; poses / indices exact, maps exact."""
    from slam_2d_lidar_scan_b200 import synthetic
    init = {"x": 0.0, "y": 0.0}
    ogp = [50, 50, init, 0.1, 2 * np.pi, 10, 360, 0.5]
    smp = [1.0, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 2]
    scene = synthetic.make_scene(seed=3, steps=6, K=360, fov=2 * np.pi, unit=0.1)

    def run(cls):
        np.random.seed(21)
        pf = cls(3, ogp, smp)
        out = []
        for fr in scene["warm"]:
            for p in pf.particles:
                p.og.updateOccupancyGrid(fr)
        for count, fr in enumerate(scene["frames"][:6], start=1):
            pf.updateParticles(fr, count)
            pf.weightUnbalanced()
            out.append(pf.poses() if hasattr(pf, "poses") else _poses(pf))
        return pf, out
    pf, got = run(S.ParticleFilter)
    ref, want = run(O.ParticleFilter)
    assert pf.geom.numSpokes == 360 and pf.geom.spokesStartIdx == 270
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    np.testing.assert_allclose(pf.weights.cpu().numpy(), [p.weight for p in ref.particles], rtol=1e-9, atol=0)
    for i in range(3):
        assert np.array_equal(pf.particles[i].og.occupancyGridTotal, ref.particles[i].og.occupancyGridTotal)
        assert np.array_equal(pf.particles[i].og.occupancyGridVisited, ref.particles[i].og.occupancyGridVisited)


def test_large_batch_properties_c3(S):
    """BASELINE config 3 size (1024 particles, 1001^2 lattices): size-independent properties instead of the oracle.
    (1) identical particles + argmax (no sampling) -> identical results in every slot, equal to the 1-particle run;
    (2) counts stay integer valued and only ever grow; (3) weights normalise to 1."""
    from slam_2d_lidar_scan_b200 import synthetic
    spec = synthetic.config("c3")
    scene = synthetic.make_scene(seed=0, steps=4, K=180, fov=np.pi, unit=0.05)
    outs = []
    for n in (1024, 1):
        np.random.seed(9)
        pf = S.ParticleFilter(n, spec["og"], spec["sm"])
        og = S.OccupancyGrid(*pf.geom.args, _geometry=pf.geom)
        for fr in scene["warm"]:
            og.updateOccupancyGrid(fr)
        pf.grids.copy_(og.device_grid.unsqueeze(0).expand_as(pf.grids))
        before = pf.grids.clone() if n == 1 else None
        # deterministic variant of the step: argmax in the coarse stage (uniforms = None)
        K = pf.geom.numSamplesPerRev
        for count, fr in enumerate(scene["frames"][:3], start=1):
            rec = pf._prepare(fr, count, n, pf._prevRaw[0], pf._prevRawHeading[0])
            pf._stage_d.copy_(pf._stage_h)
            if count > 1:
                eng, st = pf.engine, torch.cuda.current_stream().cuda_stream
                nat = S._native
                nat.check(nat.lib.slam_propose_poses(n, pf.prevMatched.data_ptr(), rec["rawTheta"], rec["prevRawTheta"], 0, 0.0,
                                                     pf.prevHeading.data_ptr(), pf.hasHeading.data_ptr(), pf._est.data_ptr(),
                                                     pf._phi.data_ptr(), pf._hasPhi.data_ptr(), pf.status.data_ptr(), st))
                n2 = eng.nOffC ** 2
                eng.match(pf.grids, n, pf._stage_d[:K], pf._est, pf._stage_d[K + n:K + n + n2], None, None, pf._matched,
                          pf._conf, pf._idx, pf.status)
                nat.check(nat.lib.slam_finish_step(n, pf._matched.data_ptr(), pf._conf.data_ptr(), pf.prevMatched.data_ptr(),
                                                   pf.prevHeading.data_ptr(), pf.hasHeading.data_ptr(), pf.weights.data_ptr(), st))
                from slam_2d_lidar_scan_b200.engine import update_grids
                update_grids(pf.geom, pf.grids, n, pf._stage_d[:K], pf._matched, pf.status)
            else:
                pf._launch(0, n, rec, pf._stage_d)
            pf._prevRaw, pf._prevRawHeading = [fr] * n, [rec["newRawHeading"]] * n
        assert int(pf.status.max().item()) == 0
        pf.weightUnbalanced()
        outs.append((pf.prevMatched.cpu().numpy(), pf._idx.cpu().numpy(), pf.weights.cpu().numpy(), pf.grids[0].cpu().numpy()))
        if n == 1024:
            assert np.all(outs[0][0] == outs[0][0][0]) and np.all(outs[0][1] == outs[0][1][0])
            assert torch.equal(pf.grids[1023], pf.grids[0]) and torch.equal(pf.grids[511], pf.grids[0])
            assert abs(outs[0][2].sum() - 1.0) < 1e-12
        else:
            g = pf.grids[0]
            assert torch.equal(g, g.round()) and bool((g >= before[0]).all())
    assert np.array_equal(outs[0][0][0], outs[1][0][0]) and np.array_equal(outs[0][1][0], outs[1][1][0])
    assert np.array_equal(outs[0][3], outs[1][3])


def test_dataset_drivers_reproduce_the_reference_loop(S, frames):
    """drivers.run_scanmatch / run_mapping are the reference's main() loops (F3): same golden poses and maps."""
    from slam_2d_lidar_scan_b200 import drivers
    data = {fr["key"]: reading(fr) for fr in frames}
    init = {"x": frames[0]["x"], "y": frames[0]["y"]}
    g = load_golden("det_c3.npz")
    og = S.OccupancyGrid(*OG_C3(init))
    sm = S.ScanMatcher(og, *SM_C3)
    poses, confs = drivers.run_scanmatch(data, og, sm, maxFrames=30)
    assert np.array_equal(poses, g["poses"])
    np.testing.assert_allclose(confs, g["confs"], rtol=CONF_RTOL, atol=0)
    g2 = load_golden("update_c3.npz")
    og2 = S.OccupancyGrid(*OG_C3(init))
    drivers.run_mapping(data, og2, maxFrames=12)
    v, t = dense_counts(int(g2["G"][0]), g2["cells"], g2["visited"], g2["total"])
    assert np.array_equal(og2.occupancyGridVisited, v) and np.array_equal(og2.occupancyGridTotal, t)
    np.random.seed(0)
    pf = S.ParticleFilter(3, [50, 50, init, 0.05, np.pi, 10, 180, 0.25], list(SM_C3))
    best, fired, b = drivers.run_fastslam(pf, data, maxFrames=22)
    g3 = load_golden("pf_c3.npz")
    assert np.array_equal(pf.poses(), g3["poses"][21]) and not fired.any()
