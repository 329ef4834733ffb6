"""ScanMatcher with the reference's surface (Utils/ScanMatcher_OGBased.py:8-176) plus the scalar odometry helpers
the reference's drivers use (:178-224).  matchScan runs the fused sm_100a kernel on the grid's device lattice."""
import json
import math

import numpy as np
import torch

from .engine import MatcherEngine, raise_for_status


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
        self.engine = _engine or MatcherEngine(og.geom, searchRadius, searchHalfRad, scanSigmaInNumGrid, moveRSigma,
                                               maxMoveDeviation, turnSigma, missMatchProbAtCoarse, coarseFactor,
                                               fineSearchHalfRad=fineSearchHalfRad)
        dev = og.geom.device
        n2 = self.engine.nOffC ** 2
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

    def matchScan(self, reading, estMovingDist, estMovingTheta, count, matchMax=True):
        """Coarse-to-fine correlative match (ScanMatcher_OGBased.py:47-79) -> (matchedReading, coarseConfidence)."""
        if count == 1:
            return reading, 1
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
        self._status.zero_()        # the kernels OR their bits in; this standalone matcher reports per call
        dbg = bufs = None
        if self.debug:
            dbg, bufs = eng.debug_buffers(1)
        eng.match(self.og.device_grid, 1, self._ranges, self._est, self._rv, tw, u, self._outPose, self._outConf,
                  self._outIdx, self._status, debug=dbg)
        pose = self._outPose.cpu().numpy()
        conf = float(self._outConf.item())
        raise_for_status(int(self._status.item()))
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
