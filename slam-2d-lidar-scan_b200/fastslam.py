"""ParticleFilter / Particle with the reference's surface (Algorithm/FastSlam.py:10-150) over a device-resident
batch of N particles, plus the FastSLAM facade (``step``) named by BASELINE.json.

State lives in torch tensors on one GPU: lattices [N][G][pitch][2] float32, poses/headings/weights float64.
``updateParticles`` is five kernel launches for all N particles (propose, priors, fused match, finish, map
update); the reference's Python loop over particles (FastSlam.py:25-27) is gone.

RNG contract: the reference draws from numpy's global legacy RandomState -- one double per particle per
``matchScan`` (coarse sampling, in particle order) and N doubles per ``resample``.  The same draws are made
here on the host, so a run seeded with ``np.random.seed`` consumes the identical stream.
"""
import math
import os

import numpy as np
import torch

from . import _native as nat
from .engine import (MatcherEngine, StepResult, SideTrigger, on_device, _NULL, raise_for_status, update_grids, plan_copy_elided, copy_lattices,
                     _stream)
from .geometry import LidarGeometry
from .grid import OccupancyGrid
from .matcher import ScanMatcher


class KernelTimer:
    """CUDA-event timing of the individual kernel launches of a step (bench.py's per-kernel breakdown).  Disabled
    (the default) it costs one attribute test per launch."""

    def __init__(self):
        self.enabled = False
        self.events = {}

    def section(self, name, dev):
        return _Section(self, name, dev) if self.enabled else _NULL

    def mean_ms(self):
        torch.cuda.synchronize()
        return {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in self.events.items() if v}


class _Section:
    def __init__(self, timer, name, dev):
        self.t, self.name, self.dev = timer, name, dev

    def __enter__(self):
        if self.t.enabled:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream(self.dev))

    def __exit__(self, *a):
        if self.t.enabled:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(torch.cuda.current_stream(self.dev))
            self.t.events.setdefault(self.name, []).append((self.e0, e1))


class Particle:
    """View of slot ``i`` of a ParticleFilter (Algorithm/FastSlam.py:64-150)."""

    def __init__(self, pf, i):
        self._pf, self._i = pf, i
        self.og = OccupancyGrid(*pf.geom.args, _geometry=pf.geom, _grids=pf.grids, _slot=i,
                                _slotmap=lambda: pf._slots_h)
        self.sm = ScanMatcher(self.og, *pf.smParameters, _engine=pf.engine)

    @property
    def weight(self):
        return float(self._pf.weights[self._i].item())

    @weight.setter
    def weight(self, w):
        self._pf.weights[self._i] = float(w)

    @property
    def xTrajectory(self):
        return [float(t[self._i, 0].item()) for t in self._pf._traj]

    @property
    def yTrajectory(self):
        return [float(t[self._i, 1].item()) for t in self._pf._traj]

    @property
    def prevMatchedReading(self):
        x, y, th = self._pf.prevMatched[self._i].cpu().tolist()
        raw = self._pf._prevRaw[self._i]
        return {'x': x, 'y': y, 'theta': th, 'range': raw['range'] if raw else None}

    @property
    def prevRawReading(self):
        return self._pf._prevRaw[self._i]

    @property
    def prevMatchedMovingTheta(self):
        if not int(self._pf.hasHeading[self._i].item()):
            return None
        return float(self._pf.prevHeading[self._i].item())

    # -- the reference's scalar helpers (FastSlam.py:77-120, 137-140), on this particle's state (host arithmetic: the
    #    batched kernels slam_propose_poses / slam_finish_step do the same for all particles inside update())
    def updateEstimatedPose(self, currentRawReading):
        from .matcher import updateEstimatedPose
        prevRaw = self.prevRawReading
        est, dist, theta, rawTheta = updateEstimatedPose(currentRawReading, self.prevMatchedReading, prevRaw,
                                                         self._pf._prevRawHeading[self._i], self.prevMatchedMovingTheta)
        self._pf._prevRawHeading[self._i] = rawTheta                                        # :102-105
        return est, dist, theta

    def getMovingTheta(self, matchedReading):
        x, y = self.xTrajectory, self.yTrajectory
        from .matcher import getMovingTheta
        return getMovingTheta(matchedReading, x, y)

    def updateTrajectory(self, matchedReading):
        pf, i = self._pf, self._i
        row = torch.zeros((pf.numParticles, 2), dtype=torch.float64, device=pf.geom.device)
        if pf._traj:
            row.copy_(pf._traj[-1])
        row[i, 0], row[i, 1] = matchedReading['x'], matchedReading['y']
        pf._traj.append(row)

    def update(self, reading, count):
        """Particle.update (FastSlam.py:122-135) for this particle only."""
        self._pf._update(self._i, self._i + 1, reading, count)
        self._pf._rawUniform = False

    def plotParticle(self):
        raise NotImplementedError("plotting is out of scope")


class ParticleFilter:
    def __init__(self, numParticles, ogParameters, smParameters, *, device=None, geometry=None, engine=None):
        self.numParticles = numParticles
        (mapX, mapY, initXY, unit, lidarFOV, lidarMaxRange, numSamplesPerRev, wallThickness) = ogParameters   # :66
        self.ogParameters, self.smParameters = list(ogParameters), list(smParameters)
        self.geom = geometry or LidarGeometry(mapX, mapY, initXY, unit, lidarFOV, numSamplesPerRev, lidarMaxRange,
                                              wallThickness, device=device)
        self.engine = engine or MatcherEngine(self.geom, *smParameters)
        self.step = 0                        # unused int attribute of the reference (FastSlam.py:15)
        self.prevMatchedReading = None       # idem (:16-18)
        self.prevRawReading = None
        self.particlesTrajectory = []
        n, dev = numParticles, self.geom.device
        f64 = dict(dtype=torch.float64, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.grids = self.geom.new_grids(n)
        # particle -> lattice table: resampling hands lattices over instead of copying them (plan_copy_elided)
        self._slots_h = np.arange(n, dtype=np.int32)
        self.slots = torch.arange(n, dtype=torch.int32, device=dev)
        self.resampleCopies = 0              # lattices physically copied by resample() so far
        # map expansion (OccupancyGrid.py:108-125): conservative host-side box around all particle poses, so that the
        # lattices can be grown BEFORE a launch would leave them, without a device read-back per step
        self._bound = None
        self._tightenBound = True            # read the true pose extremes back (one sync) before actually growing
        self.expansions = 0
        self.prevMatched = torch.zeros((n, 3), **f64)
        self.prevHeading = torch.zeros(n, **f64)
        self.hasHeading = torch.zeros(n, **i32)
        self.weights = torch.ones(n, **f64)                                  # Particle.weight = 1 (:75)
        self.status = torch.zeros(n, **i32)
        self._est = torch.zeros((n, 3), **f64)
        self._phi = torch.zeros(n, **f64)
        self._hasPhi = torch.zeros(n, **i32)
        self._matched = torch.zeros((n, 3), **f64)
        self._conf = torch.zeros(n, **f64)
        self._idx = torch.zeros((n, 6), **i32)
        n2 = self.engine.nOffC ** 2
        self._tw = torch.zeros((n, n2), **f64)
        self._rv = torch.zeros(n2, **f64)
        K = self.geom.numSamplesPerRev
        self._stage_h = torch.zeros(K + n + n2, dtype=torch.float64).pin_memory()   # ranges | uniforms | rv
        self._stage_np = self._stage_h.numpy()
        self._stage_d = torch.zeros(K + n + n2, **f64)
        self._stage_ev = torch.cuda.Event()
        self._stage_busy = False
        self._res = StepResult(dev)
        self._out = self._res.out
        self._cdf = torch.zeros(n, **f64)
        self._ridx = torch.zeros(n, **i32)
        self._traj = []
        self.keepTrajectory = True           # per-step [N][2] device copy of the matched positions
        self.kernelLaunches = 0              # kernels of this library enqueued so far
        self.matchEvents = None              # list -> (start, end) CUDA events around every match launch
        self.timer = KernelTimer()           # per-kernel CUDA events when .enabled
        # The reference dies with a TypeError (None + float, FastSlam.py:96) when a particle did not move in the step
        # before a > 0.3 m odometry step; True = carry on with a zero heading prior for that particle instead.
        self.ignoreMissingHeading = False
        self.ignoreStatusBits = 0            # further status bits not to raise on (e.g. lost particles leaving a fixed-size map)
        self.expandMaps = True               # grow the maps like the reference (False: fixed size, leaving it is an error)
        self.h2dBytes = 0
        self.d2hBytes = 0
        self._prevRaw = [None] * n
        self._prevRawHeading = [None] * n
        self._rawUniform = True
        self._particles = None
        self.lastVariance = None
        self.lastResampleIdx = None
        # Once a step's weights and poses are final (before the map update) the normalisation + trigger start on a side
        # stream (engine.SideTrigger); the sharded filter replaces the body by its all-gather + replicated normalisation.
        # Status bits raised by the map update itself (SLAM_ST_SCAN_OUTSIDE_MAP) are sticky: they surface one
        # weightUnbalanced() later.  SLAM_NO_OVERLAP=1 keeps everything on the caller's stream.
        self.overlap = not os.environ.get("SLAM_NO_OVERLAP")
        self._side = SideTrigger(dev, self._res)
        self._sideBody = self._trigger_body
        self._wn = torch.ones(n, **f64)      # side-stream copy: Particle.weight stays raw until weightUnbalanced()
        self._copyBack = None                # normalised weights whose copy into .weights is deferred (see _settle)

    # ---- reference surface
    @property
    def particles(self):
        if self._particles is None:
            self._particles = [Particle(self, i) for i in range(self.numParticles)]
        return self._particles

    def updateParticles(self, reading, count):
        """FastSlam.py:25-27 -- all particles in one batch."""
        if self._rawUniform:
            self._update(0, self.numParticles, reading, count)
        else:
            for i in range(self.numParticles):
                self._update(i, i + 1, reading, count)

    def normalizeWeights(self):
        self._copyBack = None
        if self._side.pending:      # already normalised on the side stream, into the copy
            self._side.join()
            self.weights.copy_(self._wn)
            return
        self._normalize(self.weights)

    def _trigger_body(self):
        """One launch (slam_step_trigger), normally on the side stream: normalise the weights into a COPY (the raw
        Particle.weight stays readable until weightUnbalanced() / normalizeWeights() is actually called), variance
        trigger, OR of the status words."""
        with self.timer.section("normalize_kernel", self.geom.device):
            self._res.trigger(self.weights, self._wn, self.status, self.geom.device)
        self.kernelLaunches += 1

    def weightUnbalanced(self):
        """Normalise, then the reference's variance trigger (FastSlam.py:30-41).  Synchronises (returns a bool)."""
        self._copyBack = None
        if self._side.pending:
            var, fired, bits = self._side.fetch()
            self.weights.copy_(self._wn)
        else:
            self._trigger_body()
            var, fired, bits = self._res.fetch()
            self.weights.copy_(self._wn)
        self.d2hBytes += 24
        if self.ignoreMissingHeading:
            bits &= ~nat.ST_HEADING_MISSING
        raise_for_status(bits & ~self.ignoreStatusBits)
        self.lastVariance = var
        return fired

    def resample(self):
        """np.random.choice(arange(N), N, p=weights) + deep copy of the chosen particles (FastSlam.py:50-62)."""
        n, dev = self.numParticles, self.geom.device
        u = torch.from_numpy(np.random.random_sample(n)).to(dev)
        st = _stream(dev)
        with torch.cuda.device(dev):
            nat.check(nat.lib.slam_resample_indices(n, self.weights.data_ptr(), u.data_ptr(), self._cdf.data_ptr(),
                                                    self._ridx.data_ptr(), st))
        idx = self._ridx.to(torch.int64)
        hidx = idx.cpu().tolist()
        # lattices: ownership moves with the slot table; only the extra copies of multiply-chosen particles are copied
        newSlots, copies = plan_copy_elided(hidx, self._slots_h)
        copy_lattices(self.geom, self.grids, copies)
        self._slots_h = newSlots
        self.slots.copy_(torch.from_numpy(newSlots))
        self.resampleCopies += len(copies)
        self.lastResampleCopies = len(copies)
        # small per-particle state: plain gathers in particle order
        self.prevMatched = self.prevMatched.index_select(0, idx).contiguous()
        self.prevHeading = self.prevHeading.index_select(0, idx).contiguous()
        self.hasHeading = self.hasHeading.index_select(0, idx).contiguous()
        self.weights.fill_(1.0 / n)                                      # :62 (1 / numParticles)
        self._traj = [t.index_select(0, idx) for t in self._traj]
        self._prevRaw = [self._prevRaw[i] for i in hidx]
        self._prevRawHeading = [self._prevRawHeading[i] for i in hidx]
        self.lastResampleIdx = np.asarray(hidx)

    def lattice(self, i):
        """[G][pitch][2] device view of particle i's (visited, total) lattice."""
        return self.grids[int(self._slots_h[i])]

    def load_grid(self, grid):
        """Every particle's map := ``grid`` ([G][pitch][2] float32)."""
        self.grids.copy_(grid.unsqueeze(0).expand_as(self.grids))

    # ---- device step
    def _normalize(self, w):
        with on_device(self.geom.device), self.timer.section("normalize_kernel", self.geom.device):
            nat.check(nat.lib.slam_normalize_weights(self.numParticles, w.data_ptr(), self._out.data_ptr(),
                                                     _stream(self.geom.device)))
        self.kernelLaunches += 1

    def _prepare(self, reading, count, n, prevRaw, prevRawHeading, out=None, uniforms=None):
        """Host part of Particle.update: the raw-odometry scalars (identical for every particle), the RNG draws
        and the radial prior.  Fills a pinned staging row [ranges | uniforms | rv] and returns the launch record."""
        K, N, n2 = self.geom.numSamplesPerRev, self.numParticles, self.engine.nOffC ** 2
        h = self._stage_np if out is None else out.numpy()      # numpy view of the (pinned) staging row
        h[:K] = reading['range']
        rec = dict(count=count, reading=reading, mode=0, rawTurn=0.0, newRawHeading=None)
        if count == 1:
            return rec
        raw = reading
        dx, dy = raw['x'] - prevRaw['x'], raw['y'] - prevRaw['y']
        estMovingDist = math.sqrt(dx ** 2 + dy ** 2)                                              # :82
        rawMove = math.sqrt((raw['x'] - prevRaw['x']) ** 2 + (raw['y'] - prevRaw['y']) ** 2)    # :86
        if rawMove > 0.3:                                                                         # :88-101
            rec["newRawHeading"] = math.acos(dx / rawMove) if dy > 0 else -math.acos(dx / rawMove)
            if prevRawHeading is not None:
                rec["mode"], rec["rawTurn"] = 1, rec["newRawHeading"] - prevRawHeading
        rec["rawTheta"], rec["prevRawTheta"] = raw['theta'], prevRaw['theta']
        if uniforms is None:
            uniforms = np.random.random_sample(n)                           # one draw per matchScan, particle order
        h[K:K + n] = uniforms
        h[K + N:K + N + n2] = self.engine.radial_prior(estMovingDist).reshape(-1)
        return rec

    def _launch(self, lo, hi, rec, d_stage):
        """Device part of Particle.update for particles [lo, hi): kernel launches only, no host synchronisation."""
        with on_device(self.geom.device):
            self._launch_on_device(lo, hi, rec, d_stage)

    def _launch_on_device(self, lo, hi, rec, d_stage):
        n, dev, K, N = hi - lo, self.geom.device, self.geom.numSamplesPerRev, self.numParticles
        st = _stream(dev)
        eng = self.engine
        n2 = eng.nOffC ** 2
        cut = (lambda t: t) if n == N else (lambda t: t[lo:hi])      # the whole batch: no views to build
        matched = cut(self._matched)
        status = cut(self.status)
        prevMatched, prevHeading, hasHeading = cut(self.prevMatched), cut(self.prevHeading), cut(self.hasHeading)
        est, phi, hasPhi, conf, weights = cut(self._est), cut(self._phi), cut(self._hasPhi), cut(self._conf), cut(self.weights)
        count, reading = rec["count"], rec["reading"]
        if count == 1:
            self._settle()
            # matchedReading, confidence = reading, 1 (:123-125)
            matched.copy_(torch.tensor([reading['x'], reading['y'], reading['theta']], dtype=torch.float64))
            hasHeading.zero_()
            prevMatched.copy_(matched)
        else:
            d_u = d_stage[K:K + n]
            d_rv = d_stage[K + N:K + N + n2]
            with self.timer.section("propose_kernel", dev):
                nat.check(nat.lib.slam_propose_poses(
                    n, prevMatched.data_ptr(), rec["rawTheta"], rec["prevRawTheta"], rec["mode"],
                    rec["rawTurn"], prevHeading.data_ptr(), hasHeading.data_ptr(),
                    est.data_ptr(), phi.data_ptr(), hasPhi.data_ptr(),
                    status.data_ptr(), st))
            tw = None
            if rec["mode"] == 1:
                tw = cut(self._tw)
                with self.timer.section("priors_kernel", dev):
                    nat.check(nat.lib.slam_motion_priors(n, eng.stageInfo[0]["nHalf"], eng.heading_coef,
                                                         phi.data_ptr(), hasPhi.data_ptr(),
                                                         tw.data_ptr(), st))
            if self.matchEvents is not None:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record(torch.cuda.current_stream(dev))
            with self.timer.section("match_kernel", dev):
                eng.match(self.grids, n, d_stage[:K], est, d_rv, tw, d_u, matched,
                          conf, cut(self._idx), status, slots=cut(self.slots), stream=st)
            if self.matchEvents is not None:
                ev1.record(torch.cuda.current_stream(dev))
                self.matchEvents.append((ev0, ev1))
            self._settle()         # the proposal and the match above did not need the previous step's trigger
            with self.timer.section("finish_kernel", dev):
                nat.check(nat.lib.slam_finish_step(n, matched.data_ptr(), conf.data_ptr(),
                                                   prevMatched.data_ptr(), prevHeading.data_ptr(),
                                                   hasHeading.data_ptr(), weights.data_ptr(), st))
        if self.keepTrajectory:
            if n == N:
                self._traj.append(matched[:, :2].clone())
            else:
                if not self._traj or getattr(self, "_trajCount", None) != count:
                    self._traj.append(torch.zeros((N, 2), dtype=torch.float64, device=dev))
                self._traj[-1][lo:hi] = matched[:, :2]
            self._trajCount = count
        if self.overlap and n == N:
            self._side.start(self._sideBody)
        with self.timer.section("update_kernels", dev):
            update_grids(self.geom, self.grids, n, d_stage[:K], matched, status, slots=cut(self.slots), stream=st)    # :133
        self.kernelLaunches += 3 + (0 if count == 1 else 3 + (1 if rec["mode"] == 1 else 0))

    def _settle(self):
        """Called before a launch sequence overwrites weights / poses.  The previous step's side-stream trigger must
        have read them (stream-side join, no host synchronisation), and a deferred copy-back of its normalised weights
        lands now.  Until here the next step's proposal and match ran next to that trigger: with several ranks the
        all-gather (and the skew between ranks) overlaps a whole match kernel."""
        self._side.join()
        if self._copyBack is not None:
            self.weights.copy_(self._copyBack)
            self._copyBack = None

    def _grow(self):
        """All lattices double around the map centre (geometry.grown): the filter's maps become the lattices a filter
        pre-sized to that length has.  The matcher is re-planned for the new lattice; particle views are rebuilt."""
        new, off = self.geom.grown()
        self.grids = self.geom.rehome(self.grids, new, off)
        self.geom = new
        self.engine = MatcherEngine(new, *self.smParameters)
        self.ogParameters[0], self.ogParameters[1] = new.args[0], new.args[1]
        self._particles = None
        self.expansions += 1

    def _ensure_room(self, reading, count):
        """Grow the maps if this step's search windows (est +- windowRadius, and the fine window around any coarse
        result) or scan updates could leave them."""
        if not self.expandMaps:
            return
        eng, g = self.engine, self.geom
        step = eng.searchRadius + eng.coarseStep             # a matched pose is at most this far from its proposal
        if count == 1 or self._bound is None:
            if count == 1:
                b = [reading['x'], reading['x'], reading['y'], reading['y']]      # matched = reading (:123-125)
            else:
                b = self._pose_extremes()
                b = [b[0] - step, b[1] + step, b[2] - step, b[3] + step]
        else:
            b = [self._bound[0] - step, self._bound[1] + step, self._bound[2] - step, self._bound[3] + step]
        # needed: every coarse window est +- windowRadius (:21-27) and every fine window, centred at most searchRadius
        # further out (:70).  b bounds the poses AFTER this step (proposals +- step), so b +- windowRadius holds both.
        m = 1.1 * g.lidarMaxRange + eng.searchRadius
        inside = lambda bb: self.geom.contains(bb[0] - m, bb[1] + m, bb[2] - m, bb[3] + m)
        if not inside(b) and count > 1 and self._tightenBound:
            # the running box is conservative: read the true extremes of the proposals back (one sync) before growing --
            # a map that is large enough must never be touched (growing changes the lattice's coordinate rounding)
            t = self._pose_extremes()
            sr = eng.searchRadius
            if inside([t[0] - sr, t[1] + sr, t[2] - sr, t[3] + sr]):
                b = [t[0] - step, t[1] + step, t[2] - step, t[3] + step]
                self._bound = b
                return
            b = [t[0] - step, t[1] + step, t[2] - step, t[3] + step]
        while not inside(b):
            self._grow()
        self._bound = b

    def _pose_extremes(self):
        lo, hi = self.prevMatched[:, :2].min(0).values.cpu().tolist(), self.prevMatched[:, :2].max(0).values.cpu().tolist()
        return [lo[0], hi[0], lo[1], hi[1]]

    def _update(self, lo, hi, reading, count, uniforms=None):
        """Particle.update (FastSlam.py:122-135) for particles [lo, hi): host prep, one H2D copy, launches."""
        self._ensure_room(reading, count)
        n, dev, N = hi - lo, self.geom.device, self.numParticles
        if self._stage_busy:
            self._stage_ev.synchronize()         # previous async H2D out of the pinned staging buffer has landed
            self._stage_busy = False
        rec = self._prepare(reading, count, n, self._prevRaw[lo], self._prevRawHeading[lo], uniforms=uniforms)
        self._stage_d.copy_(self._stage_h, non_blocking=True)
        self._stage_ev.record(torch.cuda.current_stream(dev))
        self._stage_busy = True
        self.h2dBytes += self._stage_h.numel() * 8
        self._launch(lo, hi, rec, self._stage_d)
        if n == N:
            self._prevRaw = [reading] * N
            self._prevRawHeading = [rec["newRawHeading"]] * N
        else:
            for i in range(lo, hi):
                self._prevRaw[i] = reading
                self._prevRawHeading[i] = rec["newRawHeading"]

    # ---- conveniences beyond the reference
    def poses(self):
        """[N][3] matched poses of the last step (host numpy)."""
        return self.prevMatched.cpu().numpy()

    def best_particle(self):
        return int(torch.argmax(self.weights).item())


class FastSLAM:
    """Facade named by BASELINE.json: ``step(reading)`` = the loop body of FastSlam.py:159-162."""

    def __init__(self, numParticles, ogParameters, smParameters, *, device=None):
        self.pf = ParticleFilter(numParticles, ogParameters, smParameters, device=device)
        self.count = 0
        self.resampled = []

    def step(self, reading):
        self.count += 1
        self.pf.updateParticles(reading, self.count)
        fired = self.pf.weightUnbalanced()
        if fired:
            self.pf.resample()
        self.resampled.append(fired)
        return fired
