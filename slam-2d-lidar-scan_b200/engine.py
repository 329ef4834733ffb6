"""Batched device engine behind the reference-shaped classes: one call = one kernel launch over N particles.

Host code only prepares constants with the reference's own Python/numpy expressions (so they are bit-identical
to what the reference would compute) and moves small buffers; all per-particle arithmetic runs in the sm_100a
kernels of csrc/ through the C ABI (include/slam2d_b200.h).
"""
import contextlib
import ctypes as C
import math

import numpy as np
import torch

from . import _native as nat


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


_NULL = contextlib.nullcontext()


def on_device(dev):
    """Context that makes ``dev`` the current CUDA device; free when it already is (the per-step host path enters it
    half a dozen times)."""
    return _NULL if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)


def gaussian_taps(sigma, truncate=4.0):
    """Taps scipy.ndimage.gaussian_filter1d would use for `sigma` (mode/truncate defaults), centre at [radius]."""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    phi = phi / phi.sum()
    return phi[::-1].copy(), radius


def raise_for_status(bits):
    if bits & nat.ST_WINDOW_OUTSIDE_MAP:
        raise IndexError("search window leaves the pre-sized map (the reference would expand it; pre-size the map)")
    if bits & nat.ST_SCAN_OUTSIDE_MAP:
        raise IndexError("scan update touches cells outside the pre-sized map")
    if bits & nat.ST_INDEX_OUT_OF_FIELD:
        raise IndexError("scan point outside the likelihood-field window")
    if bits & nat.ST_NAN_SCORE:
        raise ValueError("probabilities contain NaN")
    if bits & nat.ST_HEADING_MISSING:
        raise TypeError("unsupported operand type(s) for +: 'NoneType' and 'float' (prevMatchedMovingTheta is None)")


class MatcherEngine:
    """Device plan for ScanMatcher.matchScan (Utils/ScanMatcher_OGBased.py:47-79) on a given geometry."""

    def __init__(self, geom, searchRadius, searchHalfRad, scanSigmaInNumGrid, moveRSigma, maxMoveDeviation, turnSigma,
                 missMatchProbAtCoarse, coarseFactor, fineSearchHalfRad=None):
        self.geom = geom
        self.searchRadius, self.searchHalfRad = searchRadius, searchHalfRad
        self.scanSigmaInNumGrid, self.coarseFactor = scanSigmaInNumGrid, coarseFactor
        self.moveRSigma, self.turnSigma = moveRSigma, turnSigma
        self.missMatchProbAtCoarse, self.maxMoveDeviation = missMatchProbAtCoarse, maxMoveDeviation
        u = geom.unitGridSize
        self.coarseStep = coarseFactor * u                                   # :54
        coarseSigma = scanSigmaInNumGrid / coarseFactor                      # :55
        fineMiss = missMatchProbAtCoarse ** (2 / coarseFactor)               # :69
        self.windowRadius = 1.1 * geom.lidarMaxRange + searchRadius          # :21
        fineHalf = searchHalfRad if fineSearchHalfRad is None else fineSearchHalfRad   # :68 (extension: keyword)
        self._keep = []
        desc = nat.MatcherDesc()
        desc.windowRadius = self.windowRadius
        self.stageInfo, self.stageLog, self.stageSigma = [], [], []
        for st, ul, sigma, miss, radius, half in (
                (desc.coarse, self.coarseStep, coarseSigma, missMatchProbAtCoarse, searchRadius, searchHalfRad),
                (desc.fine, u, scanSigmaInNumGrid, fineMiss, self.coarseStep, fineHalf)):
            taps, r = gaussian_taps(sigma)
            if r < 1 or r > nat.MAX_BLUR_RADIUS:
                raise NotImplementedError("blur radius %d outside 1..%d" % (r, nat.MAX_BLUR_RADIUS))
            thetas = np.arange(-half, half + geom.angularStep, geom.angularStep)          # :114
            cosT = np.array([np.cos(t) for t in thetas])                                 # rotate() gets scalars (:169)
            sinT = np.array([np.sin(t) for t in thetas])
            st.unitLength, st.logMiss, st.blurRadius = ul, math.log(miss), r
            for i, t in enumerate(taps):
                st.blurW[i] = t
            st.nHalf = int(radius / ul)                                                   # :94
            st.nTheta = len(thetas)
            for name, arr in (("h_thetas", thetas), ("h_cos", cosT), ("h_sin", sinT)):
                arr = np.ascontiguousarray(arr, dtype=np.float64)
                self._keep.append(arr)
                setattr(st, name, arr.ctypes.data_as(nat.c_double_p))
            self.stageInfo.append(dict(unitLength=ul, nHalf=st.nHalf, thetas=thetas, radius=r))
            self.stageLog.append(math.log(miss))
            self.stageSigma.append(sigma)
        # every gather must stay inside the window: |point - pose| < maxRange, offsets <= nHalf cells
        margin = self.windowRadius - geom.lidarMaxRange
        if margin < self.stageInfo[0]["nHalf"] * self.coarseStep or margin < self.stageInfo[1]["nHalf"] * u:
            raise ValueError("search radius exceeds the window margin")
        h = C.c_void_p()
        with torch.cuda.device(geom.device):
            nat.check(nat.lib.slam_matcher_create(C.byref(geom.c), C.byref(desc), C.byref(h)))
        self.handle = h
        self.nOffC = 2 * self.stageInfo[0]["nHalf"] + 1
        self._workspace = None               # device scratch, sized for the largest batch seen so far
        self._radialDist = None
        self.zeroRv = torch.zeros(self.nOffC * self.nOffC, dtype=torch.float64, device=geom.device)

    def workspace_for(self, n):
        """Scratch for a call with n particles (one slot per CTA that gets a particle; a standalone matcher needs one)."""
        need = nat.lib.slam_matcher_workspace_bytes_n(self.handle, n)
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.geom.device)
        return self._workspace

    @property
    def workspace(self):
        """Scratch large enough for any batch (stage-level entries, slam_correlate)."""
        return self.workspace_for(1 << 30)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        try:
            if h:
                nat.lib.slam_matcher_destroy(h)
        except Exception:       # interpreter shutdown: the module globals may already be gone
            pass

    # -- priors of the coarse stage (:95-110), host side, reference expressions
    def offset_axis(self, stage=0):
        n = self.stageInfo[stage]["nHalf"]
        return np.arange(-n, n + 1)

    def radial_prior(self, estMovingDist):
        d = self._radialDist
        if d is None:                          # the offsets' distances do not depend on the step: computed once
            ax = self.offset_axis(0)
            xv, yv = np.meshgrid(ax, ax)
            ul = self.coarseStep
            d = self._radialDist = np.sqrt((xv * ul) ** 2 + (yv * ul) ** 2)
        rv = - (1 / (2 * self.moveRSigma ** 2)) * (d - estMovingDist) ** 2                # :101
        rv[np.abs(d - estMovingDist) > self.maxMoveDeviation] = -100                      # :102-103
        return rv

    def heading_prior(self, estMovingTheta):
        ax = self.offset_axis(0)
        xv, yv = np.meshgrid(ax, ax)
        if estMovingTheta is None:
            return np.zeros(xv.shape)                                                     # :110
        dist = np.sqrt(np.square(xv) + np.square(yv))
        dist[dist == 0] = 0.0001
        with np.errstate(invalid='ignore'):
            ang = np.arccos((xv * math.cos(estMovingTheta) + yv * math.sin(estMovingTheta)) / dist)   # :107
        return -1 / (2 * self.turnSigma ** 2) * np.square(ang)                            # :108

    @property
    def heading_coef(self):
        return -1 / (2 * self.turnSigma ** 2)

    def debug_buffers(self, n):
        dev = self.geom.device
        dbg = nat.MatchDebug()
        bufs = {}
        for s in range(2):
            side = nat.lib.slam_matcher_field_side(self.handle, s)
            poses = nat.lib.slam_matcher_num_poses(self.handle, s)
            bufs["prob%d" % s] = torch.zeros((n, side, side), dtype=torch.float64, device=dev)
            bufs["dims%d" % s] = torch.zeros((n, 2), dtype=torch.int32, device=dev)
            bufs["vol%d" % s] = torch.zeros((n, poses), dtype=torch.float64, device=dev)
            dbg.d_prob[s] = bufs["prob%d" % s].data_ptr()
            dbg.d_probDims[s] = bufs["dims%d" % s].data_ptr()
            dbg.d_vol[s] = bufs["vol%d" % s].data_ptr()
        return dbg, bufs

    def match(self, grids, n, d_ranges, d_estPose, d_rv, d_tw, d_uniforms, d_outPose, d_outConf, d_outIdx, d_status,
              debug=None, slots=None, stream=None):
        """slam_match_scan on the geometry's device, current stream.  All arguments are device tensors (or None).
        Status bits are OR-ed into d_status (sticky).  ``slots`` (int32 [n]): particle p reads lattice slots[p] of
        ``grids`` (all lattices of the filter) instead of lattice p."""
        dev = self.geom.device
        ws = self.workspace_for(n)
        with on_device(dev):
            nat.check(nat.lib.slam_match_scan_slots(
                self.handle, grids.data_ptr(), _ptr(slots), grids.shape[0] if slots is not None else n, n,
                d_ranges.data_ptr(), d_estPose.data_ptr(), d_rv.data_ptr(), _ptr(d_tw),
                _ptr(d_uniforms), d_outPose.data_ptr(), d_outConf.data_ptr(), d_outIdx.data_ptr(), d_status.data_ptr(),
                ws.data_ptr(), ws.numel(), C.byref(debug) if debug is not None else None,
                _stream(dev) if stream is None else stream))

    def volume_shape(self, stage):
        n = 2 * self.stageInfo[stage]["nHalf"] + 1
        return (len(self.stageInfo[stage]["thetas"]), n, n)


def update_grids(geom, grids, n, d_ranges, d_pose, d_status, slots=None, stream=None):
    st = _stream(geom.device) if stream is None else stream
    ws = geom.update_workspace(n, st)
    with on_device(geom.device):
        nat.check(nat.lib.slam_update_grid_slots(geom.c, grids.data_ptr(), _ptr(slots), n, d_ranges.data_ptr(),
                                                 d_pose.data_ptr(), d_status.data_ptr(), ws.data_ptr(), ws.numel(), st))


def plan_copy_elided(idx, slots):
    """Copy-elided resample (FastSlam.py:50-62 deep-copies every chosen particle).  idx[i] = old particle the new
    particle i is a copy of, slots[p] = lattice that holds old particle p.  The first new particle that chooses p
    simply takes over p's lattice; only the further copies of a multiply-chosen particle are physically copied, into
    the lattices of particles nobody chose.  -> (newSlots[i], [(srcLattice, dstLattice), ...]) with
    len(copies) == N - (number of distinct chosen particles)."""
    idx = np.asarray(idx, dtype=np.int64)
    slots = np.asarray(slots, dtype=np.int64)
    n = len(idx)
    newSlots = np.empty(n, dtype=np.int32)
    taken = np.zeros(n, dtype=bool)
    extra = []
    for i in range(n):
        s = idx[i]
        if not taken[s]:
            taken[s] = True
            newSlots[i] = slots[s]
        else:
            extra.append(i)
    free = slots[~taken]
    copies = []
    for i, f in zip(extra, free):
        newSlots[i] = f
        copies.append((int(slots[idx[i]]), int(f)))
    return newSlots, copies


def copy_lattices(geom, grids, copies):
    """slam_copy_lattices for a list of (srcLattice, dstLattice) pairs."""
    if not copies:
        return
    dev = geom.device
    pairs = torch.tensor(copies, dtype=torch.int32).t().contiguous().to(dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib.slam_copy_lattices(geom.c, grids.data_ptr(), len(copies), pairs[0].data_ptr(),
                                             pairs[1].data_ptr(), _stream(dev)))


class StepResult:
    """[variance, trigger] (float64) + OR of the status words (int32) in ONE 24-byte device buffer, so that
    weightUnbalanced needs a single device-to-host copy."""

    def __init__(self, dev):
        self.raw = torch.zeros(32, dtype=torch.uint8, device=dev)
        self.out = self.raw.view(torch.float64)          # [4]: variance, trigger, (status bits), -
        self.bits = self.raw.view(torch.int32)[4:5]      # low word of out[2]

    def reduce_status(self, status, dev):
        with on_device(dev):
            nat.check(nat.lib.slam_status_reduce(status.numel(), status.data_ptr(), self.bits.data_ptr(), _stream(dev)))

    def trigger(self, wIn, wOut, status, dev):
        """normalise (wIn -> wOut, sequential float64), variance + trigger -> out[0:2], OR of ``status`` -> bits: one launch."""
        with on_device(dev):
            nat.check(nat.lib.slam_step_trigger(wIn.numel(), wIn.data_ptr(), wOut.data_ptr(), status.data_ptr(),
                                                status.numel(), self.out.data_ptr(), _stream(dev)))

    def fetch(self):
        """-> (variance, fired, statusBits); synchronises."""
        h = self.raw[:24].cpu()
        return self.parse(h)

    @staticmethod
    def parse(h):
        a = h.numpy()
        v = a[:16].view(np.float64)
        return float(v[0]), bool(v[1] != 0.0), int(a[16:20].view(np.int32)[0])


class SideTrigger:
    """The end-of-step work that depends on the matched poses / weights only -- (all-gather,) sequential normalisation,
    variance trigger, status reduction, the 24-byte read-back -- enqueued on a side stream as soon as those are final,
    i.e. next to the map update instead of behind it.  ``weightUnbalanced`` then waits for the side stream alone: the
    host returns (and prepares the next scan) while the map update is still running."""

    def __init__(self, dev, res):
        self.dev, self.res = dev, res
        self.stream = torch.cuda.Stream(device=dev)
        self.ready = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.host = torch.zeros(32, dtype=torch.uint8).pin_memory()
        self.pending = False

    def start(self, body):
        """Run ``body()`` + the read-back on the side stream, ordered after everything enqueued so far."""
        self.ready.record(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.ready)
            body()
            self.host.copy_(self.res.raw, non_blocking=True)
            self.done.record(self.stream)
        self.pending = True

    def join(self):
        """Stream-side join: work enqueued from now on sees the side stream's results.  No host synchronisation."""
        if self.pending:
            torch.cuda.current_stream(self.dev).wait_event(self.done)
            self.pending = False

    def fetch(self):
        """-> (variance, fired, statusBits) of the pending side-stream work; the host waits for that work only."""
        self.done.synchronize()
        self.join()
        return StepResult.parse(self.host)
