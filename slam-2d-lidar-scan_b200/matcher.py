"""ScanMatcher with the reference's surface (Utils/ScanMatcher_OGBased.py:8-176) plus the scalar odometry helpers
the reference's drivers use (:178-224).  matchScan runs the fused sm_100a kernel on the grid's device lattice."""
import json
import math

import numpy as np
import torch

from . import _native as nat
from .engine import MatcherEngine, raise_for_status, gaussian_taps, _stream


class ScanMatcher:
    def __init__(self, og, searchRadius, searchHalfRad, scanSigmaInNumGrid, moveRSigma, maxMoveDeviation, turnSigma,
                 missMatchProbAtCoarse, coarseFactor, *, fineSearchHalfRad=None, _engine=None):
        self.og = og
        self.searchRadius = searchRadius
        self.searchHalfRad = searchHalfRad
        self.scanSigmaInNumGrid = scanSigmaInNumGrid
        self.coarseFactor = coarseFactor
        self.moveRSigma = moveRSigma
        self.turnSigma = turnSigma
        self.missMatchProbAtCoarse = missMatchProbAtCoarse
        self.maxMoveDeviation = maxMoveDeviation
        self._fineSearchHalfRad = fineSearchHalfRad
        self._shared = _engine is not None          # a particle's matcher uses its filter's engine
        self._engine = _engine or self._plan(og.geom)
        dev = og.geom.device
        n2 = self._engine.nOffC ** 2
        f64 = dict(dtype=torch.float64, device=dev)
        self._ranges = torch.zeros(og.geom.numSamplesPerRev, **f64)
        self._est = torch.zeros(3, **f64)
        self._rv = torch.zeros(n2, **f64)
        self._tw = torch.zeros(n2, **f64)
        self._u = torch.zeros(1, **f64)
        self._outPose = torch.zeros(3, **f64)
        self._outConf = torch.zeros(1, **f64)
        self._outIdx = torch.zeros(6, dtype=torch.int32, device=dev)
        self._status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.debug = False          # True: keep probSP / convTotal of the last call in self.last
        self.last = None

    def _plan(self, geom):
        return MatcherEngine(geom, self.searchRadius, self.searchHalfRad, self.scanSigmaInNumGrid, self.moveRSigma,
                             self.maxMoveDeviation, self.turnSigma, self.missMatchProbAtCoarse, self.coarseFactor,
                             fineSearchHalfRad=self._fineSearchHalfRad)

    @property
    def engine(self):
        """Device plan for the map's CURRENT lattice (re-planned after the map has grown)."""
        if not self._shared and self._engine.geom is not self.og.geom:
            self._engine = self._plan(self.og.geom)
        return self._engine

    def _cover_search(self, x, y):
        """frameSearchSpace grows the map until it holds the search window est +- (1.1 maxRange + searchRadius)
        (:21-27): the same test, before the fused launch (a map that is large enough is never touched)."""
        if not self._shared:
            m = 1.1 * self.og.lidarMaxRange + self.searchRadius
            self.og._cover(x - m, x + m, y - m, y + m)

    def matchScan(self, reading, estMovingDist, estMovingTheta, count, matchMax=True):
        """Coarse-to-fine correlative match (ScanMatcher_OGBased.py:47-79) -> (matchedReading, coarseConfidence)."""
        if count == 1:
            return reading, 1
        self._cover_search(reading['x'], reading['y'])
        eng = self.engine
        rng = np.asarray(reading['range'], dtype=np.float64)
        self._ranges.copy_(torch.from_numpy(rng))
        self._est.copy_(torch.tensor([reading['x'], reading['y'], reading['theta']], dtype=torch.float64))
        self._rv.copy_(torch.from_numpy(eng.radial_prior(estMovingDist).reshape(-1)))
        tw = None
        if estMovingTheta is not None:
            self._tw.copy_(torch.from_numpy(eng.heading_prior(estMovingTheta).reshape(-1)))
            tw = self._tw
        u = None
        if not matchMax:
            # np.random.choice(arange(n), 1, p=...) draws exactly one double from the legacy global RandomState
            self._u.copy_(torch.from_numpy(np.random.random_sample(1)))
            u = self._u
        while True:
            self._status.zero_()        # the kernels OR their bits in; this standalone matcher reports per call
            dbg = bufs = None
            if self.debug:
                dbg, bufs = eng.debug_buffers(1)
            eng.match(self.og.device_grid, 1, self._ranges, self._est, self._rv, tw, u, self._outPose, self._outConf,
                      self._outIdx, self._status, debug=dbg)
            pose = self._outPose.cpu().numpy()
            conf = float(self._outConf.item())
            bits = int(self._status.item())
            if (bits & nat.ST_WINDOW_OUTSIDE_MAP) and not self._shared:
                # the FINE window (around the coarse result) left the map: the reference grows the map in its second
                # frameSearchSpace call (:70); here the map grows and the fused match is repeated on the new lattice
                self.og.expandOccupancyGrid(0)
                eng = self.engine
                continue
            break
        raise_for_status(bits)
        idx = self._outIdx.cpu().numpy()
        self.lastIdx = (tuple(int(v) for v in idx[:3]), tuple(int(v) for v in idx[3:]))
        if bufs is not None:
            self.last = {}
            for s, tag in enumerate(("coarse", "fine")):
                rows, cols = (int(v) for v in bufs["dims%d" % s][0].cpu())
                self.last[tag + "_prob"] = bufs["prob%d" % s][0, :rows, :cols].cpu().numpy()
                self.last[tag + "_vol"] = bufs["vol%d" % s][0].cpu().numpy().reshape(eng.volume_shape(s))
        matched = {"x": float(pose[0]), "y": float(pose[1]), "theta": float(pose[2]), "range": reading['range']}
        return matched, conf

    # ---- the reference's public stage methods (ScanMatcher_OGBased.py:20-45, 81-176).  matchScan fuses them on the
    #      device; these run ONE stage through the stage-level C entries (slam_field_build / slam_correlate /
    #      slam_blur_clamp) and return the reference's shapes.
    def _stage_of(self, unitLength):
        eng = self.engine
        for s in (0, 1):
            if unitLength == eng.stageInfo[s]["unitLength"]:
                return s
        raise NotImplementedError("unitLength %r is neither the coarse step %r nor the map unit %r this matcher was "
                                  "planned for" % (unitLength, eng.stageInfo[0]["unitLength"], eng.stageInfo[1]["unitLength"]))

    def frameSearchSpace(self, estimatedX, estimatedY, unitLength, sigma, missMatchProbAtCoarse):
        """:20-39 -> (xRangeList, yRangeList, probSP).  sigma / missMatchProb must be the stage's own (they are baked
        into the device plan: scipy taps, first-pass table, log(missProb))."""
        eng = self.engine
        s = self._stage_of(unitLength)
        if gaussian_taps(sigma)[1] != eng.stageInfo[s]["radius"] or abs(math.log(missMatchProbAtCoarse) - eng.stageLog[s]) > 0 \
                or sigma != eng.stageSigma[s]:
            raise NotImplementedError("frameSearchSpace: sigma / missMatchProb differ from the planned stage")
        maxScanRadius = 1.1 * self.og.lidarMaxRange + self.searchRadius                      # :21
        xRangeList = [estimatedX - maxScanRadius, estimatedX + maxScanRadius]
        yRangeList = [estimatedY - maxScanRadius, estimatedY + maxScanRadius]
        self.og.checkAndExapndOG(xRangeList, yRangeList)                                     # :27
        dev = self.og.geom.device
        side = nat.lib.slam_matcher_field_side(eng.handle, s)
        prob = torch.zeros((side, side), dtype=torch.float64, device=dev)
        dims = torch.zeros(2, dtype=torch.int32, device=dev)
        self._est.copy_(torch.tensor([estimatedX, estimatedY, 0.0], dtype=torch.float64))
        self._status.zero_()
        ws = eng.workspace_for(1)
        with torch.cuda.device(dev):
            nat.check(nat.lib.slam_field_build(eng.handle, s, self.og.device_grid.data_ptr(), 1, self._est.data_ptr(),
                                               prob.data_ptr(), dims.data_ptr(), self._status.data_ptr(),
                                               ws.data_ptr(), ws.numel(), _stream(dev)))
        raise_for_status(int(self._status.item()))
        rows, cols = (int(v) for v in dims.cpu())
        return xRangeList, yRangeList, prob[:rows, :cols].cpu().numpy()

    def generateProbSearchSpace(self, searchSpace, sigma):
        """:41-45 on an arbitrary array: gaussian_filter (scipy's operation order), min, clamp -- on the device."""
        dev = self.og.geom.device
        a = torch.from_numpy(np.ascontiguousarray(searchSpace, dtype=np.float64)).to(dev)
        taps, r = gaussian_taps(sigma)
        w = torch.from_numpy(np.ascontiguousarray(taps)).to(dev)
        tmp, out = torch.empty_like(a), torch.empty_like(a)
        with torch.cuda.device(dev):
            nat.check(nat.lib.slam_blur_clamp(a.data_ptr(), a.shape[0], a.shape[1], w.data_ptr(), r, tmp.data_ptr(),
                                              out.data_ptr(), _stream(dev)))
        return out.cpu().numpy()

    def covertMeasureToXY(self, estimatedX, estimatedY, estimatedTheta, rMeasure):
        """:81-89 (host numpy: a public helper of the plots; the kernels project the beams themselves)."""
        rMeasure = np.asarray(rMeasure)
        rads = np.linspace(estimatedTheta - self.og.lidarFOV / 2, estimatedTheta + self.og.lidarFOV / 2,
                           num=self.og.numSamplesPerRev)
        range_idx = rMeasure < self.og.lidarMaxRange
        rMeasureInRange = rMeasure[range_idx]
        rads = rads[range_idx]
        return estimatedX + np.cos(rads) * rMeasureInRange, estimatedY + np.sin(rads) * rMeasureInRange

    def rotate(self, origin, point, angle):
        """:162-171"""
        ox, oy = origin
        px, py = point
        qx = ox + np.cos(angle) * (px - ox) - np.sin(angle) * (py - oy)
        qy = oy + np.sin(angle) * (px - ox) + np.cos(angle) * (py - oy)
        return qx, qy

    def convertXYToSearchSpaceIdx(self, px, py, beginX, beginY, unitLength):
        """:173-176"""
        xIdx = (((np.asarray(px) - beginX) / unitLength)).astype(int)
        yIdx = (((np.asarray(py) - beginY) / unitLength)).astype(int)
        return xIdx, yIdx

    def searchToMatch(self, probSP, estimatedX, estimatedY, estimatedTheta, rMeasure, xRangeList, yRangeList,
                      searchRadius, searchHalfRad, unitLength, estMovingDist, estMovingTheta, fineSearch=False,
                      matchMax=True):
        """:91-151 against the caller's probSP -> (matchedPx, matchedPy, matchedReading, convTotal, confidence)."""
        eng = self.engine
        s = self._stage_of(unitLength)
        info = eng.stageInfo[s]
        n = int(searchRadius / unitLength)                                                   # :94
        thetaRange = np.arange(-searchHalfRad, searchHalfRad + self.og.angularStep, self.og.angularStep)   # :114
        if n != info["nHalf"] or len(thetaRange) != len(info["thetas"]) or not np.array_equal(thetaRange, info["thetas"]):
            raise NotImplementedError("searchToMatch: search radius / half angle differ from the planned stage")
        dev = self.og.geom.device
        side = nat.lib.slam_matcher_field_side(eng.handle, s)
        probSP = np.asarray(probSP, dtype=np.float64)
        rows, cols = probSP.shape
        if rows > side or cols > side:
            raise IndexError("probSP larger than the planned window")
        prob = torch.zeros((side, side), dtype=torch.float64, device=dev)
        prob[:rows, :cols] = torch.from_numpy(np.ascontiguousarray(probSP)).to(dev)
        f64 = dict(dtype=torch.float64, device=dev)
        dims = torch.tensor([rows, cols], dtype=torch.int32, device=dev)
        rMeasure = np.asarray(rMeasure, dtype=np.float64)
        self._ranges.copy_(torch.from_numpy(rMeasure))
        centre = torch.tensor([estimatedX, estimatedY, estimatedTheta], **f64)
        origin = torch.tensor([xRangeList[0], yRangeList[0]], **f64)
        nOff = 2 * n + 1
        rv = tw = None
        if not fineSearch:                                                                   # :98-110
            ax = np.arange(-n, n + 1)
            xv, yv = np.meshgrid(ax, ax)
            d = np.sqrt((xv * unitLength) ** 2 + (yv * unitLength) ** 2)
            rvh = - (1 / (2 * self.moveRSigma ** 2)) * (d - estMovingDist) ** 2
            rvh[np.abs(d - estMovingDist) > self.maxMoveDeviation] = -100
            rv = torch.from_numpy(rvh.reshape(-1)).to(dev)
            if estMovingTheta is not None:
                dist = np.sqrt(np.square(xv) + np.square(yv))
                dist[dist == 0] = 0.0001
                with np.errstate(invalid='ignore'):
                    ang = np.arccos((xv * math.cos(estMovingTheta) + yv * math.sin(estMovingTheta)) / dist)
                tw = torch.from_numpy((-1 / (2 * self.turnSigma ** 2) * np.square(ang)).reshape(-1)).to(dev)
        u = None
        if not matchMax:
            u = torch.from_numpy(np.random.random_sample(1)).to(dev)                         # one draw, like np.random.choice
        poses = nat.lib.slam_matcher_num_poses(eng.handle, s)
        vol = torch.zeros(poses, **f64)
        outIdx = torch.zeros(3, dtype=torch.int32, device=dev)
        self._status.zero_()
        ws = eng.workspace_for(1)
        with torch.cuda.device(dev):
            nat.check(nat.lib.slam_correlate(
                eng.handle, s, 1, prob.data_ptr(), dims.data_ptr(), self._ranges.data_ptr(), centre.data_ptr(),
                origin.data_ptr(), 0 if rv is None else rv.data_ptr(), 0 if tw is None else tw.data_ptr(),
                0 if u is None else u.data_ptr(), vol.data_ptr(), outIdx.data_ptr(), self._outConf.data_ptr(),
                self._status.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)))
        confidence = float(self._outConf.item())
        raise_for_status(int(self._status.item()))
        it, iy, ix = (int(v) for v in outIdx.cpu())
        convTotal = vol.cpu().numpy().reshape(len(thetaRange), nOff, nOff)
        ax = np.arange(-n, n + 1)
        dx, dy, dtheta = ax[ix] * unitLength, ax[iy] * unitLength, thetaRange[it]             # :142
        matchedReading = {"x": estimatedX + dx, "y": estimatedY + dy, "theta": estimatedTheta + dtheta, "range": rMeasure}
        px, py = self.covertMeasureToXY(estimatedX, estimatedY, estimatedTheta, rMeasure)
        matchedPx, matchedPy = self.rotate((estimatedX, estimatedY), (px, py), dtheta)
        return matchedPx + dx, matchedPy + dy, matchedReading, convTotal, confidence

    def plotMatchOverlay(self, *a, **k):
        raise NotImplementedError("plotting is out of scope")


# ---- scalar host helpers of the reference's drivers (ScanMatcher_OGBased.py:178-224 == FastSlam.py:77-120)
def _signedHeading(dx, dy, d):
    return math.acos(dx / d) if dy > 0 else -math.acos(dx / d)


def updateEstimatedPose(currentRawReading, prevMatchedReading, prevRawReading, prevRawMovingTheta,
                        prevMatchedMovingTheta):
    estimatedTheta = prevMatchedReading['theta'] + currentRawReading['theta'] - prevRawReading['theta']
    estimatedReading = {'x': prevMatchedReading['x'], 'y': prevMatchedReading['y'], 'theta': estimatedTheta,
                        'range': currentRawReading['range']}
    dx = currentRawReading['x'] - prevRawReading['x']
    dy = currentRawReading['y'] - prevRawReading['y']
    estMovingDist = math.sqrt(dx ** 2 + dy ** 2)
    rawMove = math.sqrt((currentRawReading['x'] - prevRawReading['x']) ** 2 +
                        (currentRawReading['y'] - prevRawReading['y']) ** 2)
    rawMovingTheta = estMovingTheta = None
    if rawMove > 0.3:
        rawMovingTheta = _signedHeading(dx, dy, rawMove)
        if prevRawMovingTheta is not None:
            estMovingTheta = prevMatchedMovingTheta + (rawMovingTheta - prevRawMovingTheta)
    return estimatedReading, estMovingDist, estMovingTheta, rawMovingTheta


def updateTrajectory(matchedReading, xTrajectory, yTrajectory):
    xTrajectory.append(matchedReading['x'])
    yTrajectory.append(matchedReading['y'])


def getMovingTheta(matchedReading, xTrajectory, yTrajectory):
    xMove, yMove = matchedReading['x'] - xTrajectory[-1], matchedReading['y'] - yTrajectory[-1]
    move = math.sqrt(xMove ** 2 + yMove ** 2)
    return _signedHeading(xMove, yMove, move) if move != 0 else None


def readJson(jsonFile):
    with open(jsonFile, 'r') as f:
        return json.load(f)['map']
