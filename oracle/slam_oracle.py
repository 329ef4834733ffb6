"""CPU oracle: numpy restatement of the reference's scan-match / FastSLAM hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the product package
(``slam-2d-lidar-scan_b200``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
CPU-baseline legs of ``bench.py`` may import this file, and only as the checker /
the reported CPU baseline.

Parity pin: the reference ships no tests or golden vectors, so
this oracle is pinned against outputs of the reference itself, run in the build
container by ``tests/golden/make_golden.py`` (fixtures committed under
``tests/golden/``) -- see ``tests/test_oracle_golden.py``.

What it restates (all paths relative to /root/reference):
  GridGeometry            Utils/OccupancyGrid.py:7-57     (lattice, sector tables)
  map_index               Utils/OccupancyGrid.py:102-106
  update_counts           Utils/OccupancyGrid.py:127-152
  occupancy_scatter       Utils/ScanMatcher_OGBased.py:20-37
  likelihood_field        Utils/ScanMatcher_OGBased.py:38-45
  beam_endpoints          Utils/ScanMatcher_OGBased.py:81-89
  motion_priors           Utils/ScanMatcher_OGBased.py:94-110
  correlation_volume      Utils/ScanMatcher_OGBased.py:112-132,162-176
  choose_pose             Utils/ScanMatcher_OGBased.py:133-144
  match_scan              Utils/ScanMatcher_OGBased.py:47-79
  propose_pose / heading  Algorithm/FastSlam.py:77-120 (= ScanMatcher_OGBased.py:178-224)
  normalize / unbalanced / resample_indices   Algorithm/FastSlam.py:30-62

Deliberate deviations: the map never expands -- a search window or
a scan that leaves the pre-sized lattice raises ``IndexError``; maps are square.

The structure is functional (state = two count arrays per particle + small pose
records) rather than the reference's object graph; the thin classes at the bottom
only re-create the reference's call surface so parity tests read like its drivers.
"""
import math
import numpy as np
from scipy.ndimage import gaussian_filter

from .primitives import legacy_choice_index


# --------------------------------------------------------------------------- geometry
class GridGeometry:
    """Pose-independent lattice + lidar sector tables (Utils/OccupancyGrid.py:7-57).

    One instance is shared by every particle (the reference rebuilds it per particle).
    """

    def __init__(self, mapXLength, mapYLength, initXY, unitGridSize, lidarFOV,
                 numSamplesPerRev, lidarMaxRange, wallThickness):
        xNum = int(mapXLength / unitGridSize)
        yNum = int(mapYLength / unitGridSize)
        if xNum != yNum:
            raise NotImplementedError("non-square maps are a reference defect (SURVEY A.8); not modelled")
        u = unitGridSize
        self.unit = u
        self.G = xNum + 1
        # OccupancyGrid.py:10-11 (y re-uses xNum for its extent)
        self.gridX = np.linspace(-xNum * u / 2, xNum * u / 2, num=xNum + 1) + initXY['x']
        self.gridY = np.linspace(-xNum * u / 2, xNum * u / 2, num=yNum + 1) + initXY['y']
        self.mapXLim = [self.gridX[0], self.gridX[-1]]
        self.mapYLim = [self.gridY[0], self.gridY[-1]]
        self.fov = lidarFOV
        self.maxRange = lidarMaxRange
        self.wall = wallThickness
        self.K = numSamplesPerRev
        self.angularStep = lidarFOV / numSamplesPerRev                          # :22
        self.numSpokes = int(np.rint(2 * np.pi / self.angularStep))             # :23
        self.spokesStartIdx = int(((self.numSpokes / 2 - self.K) / 2) % self.numSpokes)  # :30
        self._sector_tables()

    def _sector_tables(self):
        """OccupancyGrid.py:32-45: bearing sector + radius of every lidar-local cell."""
        h = int(self.maxRange / self.unit)
        L = 2 * h + 1
        ax = np.linspace(-self.maxRange, self.maxRange, L)
        lx, ly = np.meshgrid(ax, ax)
        sector = np.zeros((L, L))
        right = slice(h + 1, L)
        with np.errstate(divide='ignore', invalid='ignore'):
            sector[:, right] = np.rint((np.pi / 2 + np.arctan(ly[:, right] / lx[:, right]))
                                       / np.pi / 2 * self.numSpokes - 0.5).astype(int)
        # left half = point-mirrored right half + half a turn; centre column split by sign of y
        sector[:, 0:h] = np.fliplr(np.flipud(sector))[:, 0:h] + int(self.numSpokes / 2)
        sector[h + 1:L, h] = int(self.numSpokes / 2)
        self.L = L
        self.localAxis = ax                                  # lidar-local x (and y) coordinate per column (row)
        self.sector = sector.astype(np.int32)                # values 0..numSpokes-1
        self.radius = np.sqrt(lx ** 2 + ly ** 2)

    def new_counts(self):
        """OccupancyGrid.py:13-14: (visited, total) initialised to (1, 2)."""
        return np.ones((self.G, self.G)), 2 * np.ones((self.G, self.G))

    def map_index(self, x, y):
        """OccupancyGrid.py:102-106 (round-half-even)."""
        xi = np.rint((np.asarray(x) - self.mapXLim[0]) / self.unit).astype(int)
        yi = np.rint((np.asarray(y) - self.mapYLim[0]) / self.unit).astype(int)
        return xi, yi


# --------------------------------------------------------------------------- map update
def update_counts(geom, visited, total, x, y, theta, ranges):
    """OccupancyGrid.updateOccupancyGrid (OccupancyGrid.py:127-152), all beams at once.

    Per beam i the reference marks sector cells with r < range-wall/2 as seen-empty
    (total += 1, only if range < maxRange) and range-wall/2 < r < range+wall/2 as hit
    (visited += 2, total += 2); numpy's fancy ``+=`` applies once per distinct map cell
    per statement, i.e. union within a beam, accumulation across beams.
    """
    ranges = np.asarray(ranges, dtype=np.float64)
    off = int(np.rint(theta / (2 * np.pi) * geom.numSpokes))                         # :131
    beam = (geom.sector - geom.spokesStartIdx - off) % geom.numSpokes                # inverse of :134
    inFan = beam < geom.K
    rm = np.where(inFan, ranges[np.minimum(beam, geom.K - 1)], np.nan)
    half = geom.wall / 2
    empty = inFan & (rm < geom.maxRange) & (geom.radius < rm - half)                 # :138-139
    hit = inFan & (geom.radius > rm - half) & (geom.radius < rm + half)              # :142-143
    ly, lx = np.nonzero(empty | hit)
    xi, yi = geom.map_index(x + geom.localAxis[lx], y + geom.localAxis[ly])          # :144-145
    if xi.size and (xi.min() < 0 or yi.min() < 0 or xi.max() >= geom.G or yi.max() >= geom.G):
        raise IndexError("scan leaves the pre-sized map (expansion is out of scope)")
    cell = yi * geom.G + xi
    b = beam[ly, lx]
    for mask, dv, dt in ((empty[ly, lx], 0.0, 1.0), (hit[ly, lx], 2.0, 2.0)):
        # one application per distinct (beam, map cell)
        pairs = np.unique(b[mask].astype(np.int64) * (geom.G * geom.G) + cell[mask])
        cells = pairs % (geom.G * geom.G)
        np.add.at(total.reshape(-1), cells, dt)
        if dv:
            np.add.at(visited.reshape(-1), cells, dv)


# --------------------------------------------------------------------------- likelihood field
def occupancy_scatter(geom, visited, total, cx, cy, unitLength, windowRadius, missProb):
    """ScanMatcher.frameSearchSpace up to the blur (ScanMatcher_OGBased.py:20-37)."""
    xr = [cx - windowRadius, cx + windowRadius]
    yr = [cy - windowRadius, cy + windowRadius]
    nx = int((xr[1] - xr[0]) / unitLength)
    ny = int((yr[1] - yr[0]) / unitLength)
    space = math.log(missProb) * np.ones((ny + 1, nx + 1))
    if xr[0] < geom.mapXLim[0] or xr[1] > geom.mapXLim[1] or yr[0] < geom.mapYLim[0] or yr[1] > geom.mapYLim[1]:
        # the reference would grow the map here (OccupancyGrid.py:108-125)
        raise IndexError("search window leaves the pre-sized map (expansion is out of scope)")
    xi, yi = geom.map_index(xr, yr)
    v = visited[yi[0]:yi[1], xi[0]:xi[1]]
    t = total[yi[0]:yi[1], xi[0]:xi[1]]
    occ = v / t > 0.5
    # 1-D index maps (rows of OccupancyGridX / columns of OccupancyGridY are identical)
    col = ((geom.gridX[xi[0]:xi[1]] - xr[0]) / unitLength).astype(int)
    row = ((geom.gridY[yi[0]:yi[1]] - yr[0]) / unitLength).astype(int)
    oy, ox = np.nonzero(occ)
    space[row[oy], col[ox]] = 0
    return xr[0], yr[0], space


def likelihood_field(space, sigma):
    """ScanMatcher.generateProbSearchSpace (ScanMatcher_OGBased.py:41-45)."""
    prob = gaussian_filter(space, sigma=sigma)
    lo = prob.min()
    prob[prob > 0.5 * lo] = 0
    return prob


# --------------------------------------------------------------------------- correlative search
def beam_endpoints(geom, x, y, theta, ranges):
    """ScanMatcher.covertMeasureToXY (ScanMatcher_OGBased.py:81-89)."""
    ang = np.linspace(theta - geom.fov / 2, theta + geom.fov / 2, num=geom.K)
    keep = ranges < geom.maxRange
    r = ranges[keep]
    ang = ang[keep]
    return x + np.cos(ang) * r, y + np.sin(ang) * r


def offset_axis(searchRadius, unitLength):
    n = int(searchRadius / unitLength)                                   # :94
    return np.arange(-n, n + 1)


def motion_priors(axis, unitLength, estMovingDist, estMovingTheta, moveRSigma, maxMoveDeviation, turnSigma):
    """Coarse-stage radial and heading priors (ScanMatcher_OGBased.py:97-110).  [ny, nx]."""
    xv, yv = np.meshgrid(axis, axis)
    d = np.sqrt((xv * unitLength) ** 2 + (yv * unitLength) ** 2)
    rv = - (1 / (2 * moveRSigma ** 2)) * (d - estMovingDist) ** 2
    rv[np.abs(d - estMovingDist) > maxMoveDeviation] = -100
    if estMovingTheta is None:
        return rv, np.zeros(xv.shape)
    dist = np.sqrt(np.square(xv) + np.square(yv))
    dist[dist == 0] = 0.0001
    with np.errstate(invalid='ignore'):
        ang = np.arccos((xv * math.cos(estMovingTheta) + yv * math.sin(estMovingTheta)) / dist)
    return rv, -1 / (2 * turnSigma ** 2) * np.square(ang)


def theta_offsets(geom, searchHalfRad):
    return np.arange(-searchHalfRad, searchHalfRad + geom.angularStep, geom.angularStep)   # :114


def correlation_volume(prob, px, py, ox, oy, beginX, beginY, unitLength, thetas, axis, rv, tw):
    """Score every (theta, dy, dx) hypothesis (ScanMatcher_OGBased.py:115-132)."""
    n = axis.shape[0]
    vol = np.zeros((len(thetas), n, n))
    dy = axis.reshape(n, 1, 1)
    dx = axis.reshape(1, n, 1)
    for i, th in enumerate(thetas):
        c, s = np.cos(th), np.sin(th)
        qx = ox + c * (px - ox) - s * (py - oy)                              # rotate :169-170
        qy = oy + s * (px - ox) + c * (py - oy)
        xi = ((qx - beginX) / unitLength).astype(int)                       # :174-175
        yi = ((qy - beginY) / unitLength).astype(int)
        pts = np.unique(np.column_stack((xi, yi)), axis=0)                   # :120
        gathered = prob[pts[:, 1].reshape(1, 1, -1) + dy, pts[:, 0].reshape(1, 1, -1) + dx]
        vol[i] = np.sum(gathered, axis=2) + rv + tw
    return vol


def choose_pose(vol, matchMax, uniform=None):
    """argmax / softmax sample + confidence (ScanMatcher_OGBased.py:133-141).

    ``uniform``: the one double np.random.choice would draw (legacy RandomState)."""
    if matchMax:
        flat = int(vol.argmax())
    else:
        e = np.exp(vol.reshape(-1))
        p = e / e.sum()
        if np.isnan(p).any():
            raise ValueError("probabilities contain NaN")
        if uniform is None:
            uniform = np.random.random_sample(1)
        flat = int(legacy_choice_index(p, np.atleast_1d(uniform))[0])
    conf = np.sum(np.exp(vol))
    return np.unravel_index(flat, vol.shape), conf


class MatcherParams:
    """ScanMatcher.__init__ arguments (ScanMatcher_OGBased.py:9-18)."""

    def __init__(self, searchRadius, searchHalfRad, scanSigmaInNumGrid, moveRSigma, maxMoveDeviation,
                 turnSigma, missMatchProbAtCoarse, coarseFactor):
        self.searchRadius = searchRadius
        self.searchHalfRad = searchHalfRad
        self.sigma = scanSigmaInNumGrid
        self.moveRSigma = moveRSigma
        self.maxMoveDeviation = maxMoveDeviation
        self.turnSigma = turnSigma
        self.missProb = missMatchProbAtCoarse
        self.coarseFactor = coarseFactor


def search_stage(geom, mp, visited, total, x, y, theta, ranges, searchRadius, unitLength, sigma, missProb,
                 estMovingDist, estMovingTheta, fine, matchMax, uniform=None, trace=None):
    """frameSearchSpace + searchToMatch for one stage (ScanMatcher_OGBased.py:57-60 / :70-73)."""
    windowRadius = 1.1 * geom.maxRange + mp.searchRadius        # :21 -- always the constructor radius
    bx, by, space = occupancy_scatter(geom, visited, total, x, y, unitLength, windowRadius, missProb)
    prob = likelihood_field(space, sigma)
    px, py = beam_endpoints(geom, x, y, theta, ranges)
    axis = offset_axis(searchRadius, unitLength)
    if fine:
        rv = tw = np.zeros((axis.shape[0], axis.shape[0]))       # :99
    else:
        rv, tw = motion_priors(axis, unitLength, estMovingDist, estMovingTheta, mp.moveRSigma,
                               mp.maxMoveDeviation, mp.turnSigma)
    thetas = theta_offsets(geom, mp.searchHalfRad)
    vol = correlation_volume(prob, px, py, x, y, bx, by, unitLength, thetas, axis, rv, tw)
    (it, iy, ix), conf = choose_pose(vol, matchMax, uniform)
    dx, dy, dth = axis[ix] * unitLength, axis[iy] * unitLength, thetas[it]      # :142
    if trace is not None:
        trace.append(dict(prob=prob, vol=vol, idx=(int(it), int(iy), int(ix)), conf=float(conf)))
    return x + dx, y + dy, theta + dth, conf, (int(it), int(iy), int(ix))


def match_scan(geom, mp, visited, total, x, y, theta, ranges, estMovingDist, estMovingTheta,
               matchMax=True, uniform=None, trace=None):
    """ScanMatcher.matchScan for count >= 2 (ScanMatcher_OGBased.py:53-79).

    Returns (x, y, theta, coarseConfidence, coarseIdx, fineIdx)."""
    ranges = np.asarray(ranges, dtype=np.float64)
    cstep = mp.coarseFactor * geom.unit
    cx, cy, cth, conf, cidx = search_stage(
        geom, mp, visited, total, x, y, theta, ranges, mp.searchRadius, cstep, mp.sigma / mp.coarseFactor,
        mp.missProb, estMovingDist, estMovingTheta, fine=False, matchMax=matchMax, uniform=uniform, trace=trace)
    fx, fy, fth, _, fidx = search_stage(
        geom, mp, visited, total, cx, cy, cth, ranges, cstep, geom.unit, mp.sigma,
        mp.missProb ** (2 / mp.coarseFactor), estMovingDist, estMovingTheta, fine=True, matchMax=True, trace=trace)
    return fx, fy, fth, conf, cidx, fidx


# --------------------------------------------------------------------------- odometry proposal
def _heading(dx, dy, d):
    return math.acos(dx / d) if dy > 0 else -math.acos(dx / d)


def propose_pose(raw, prevMatched, prevRaw, prevRawMovingTheta, prevMatchedMovingTheta):
    """updateEstimatedPose (FastSlam.py:77-106 / ScanMatcher_OGBased.py:178-205).

    Returns (estX, estY, estTheta, estMovingDist, estMovingTheta|None, rawMovingTheta|None)."""
    estTheta = prevMatched['theta'] + raw['theta'] - prevRaw['theta']
    dx, dy = raw['x'] - prevRaw['x'], raw['y'] - prevRaw['y']
    dist = math.sqrt(dx ** 2 + dy ** 2)
    rawMove = math.sqrt((raw['x'] - prevRaw['x']) ** 2 + (raw['y'] - prevRaw['y']) ** 2)
    rawMovingTheta = estMovingTheta = None
    if rawMove > 0.3:
        rawMovingTheta = _heading(dx, dy, rawMove)
        if prevRawMovingTheta is not None:
            estMovingTheta = prevMatchedMovingTheta + (rawMovingTheta - prevRawMovingTheta)
    return prevMatched['x'], prevMatched['y'], estTheta, dist, estMovingTheta, rawMovingTheta


def moving_heading(x, y, prevX, prevY):
    """getMovingTheta (FastSlam.py:108-120)."""
    mx, my = x - prevX, y - prevY
    d = math.sqrt(mx ** 2 + my ** 2)
    return _heading(mx, my, d) if d != 0 else None


# --------------------------------------------------------------------------- particle weights
def normalize(weights):
    """ParticleFilter.normalizeWeights: sequential float64 sum, then divide (FastSlam.py:43-48)."""
    s = 0
    for w in weights:
        s += w
    return [w / s for w in weights]


def unbalanced(weights):
    """ParticleFilter.weightUnbalanced's trigger on already-normalised weights (FastSlam.py:32-41)."""
    n = len(weights)
    var = 0
    for w in weights:
        var += (w - 1 / n) ** 2
    return bool(var > ((n - 1) / n) ** 2 + (n - 1.000000000000001) * (1 / n) ** 2), var


def resample_indices(weights, uniforms=None):
    """np.random.choice(arange(N), N, p=weights) (FastSlam.py:59) via the legacy CDF inversion."""
    w = np.asarray(weights, dtype=np.float64)
    if uniforms is None:
        uniforms = np.random.random_sample(w.shape[0])
    return legacy_choice_index(w, uniforms)


# --------------------------------------------------------------------------- reference-shaped surface
class OccupancyGrid:
    """Call surface of Utils/OccupancyGrid.py:6 over the functional core."""

    def __init__(self, mapXLength, mapYLength, initXY, unitGridSize, lidarFOV, numSamplesPerRev, lidarMaxRange,
                 wallThickness, geometry=None):
        self.geom = geometry or GridGeometry(mapXLength, mapYLength, initXY, unitGridSize, lidarFOV,
                                             numSamplesPerRev, lidarMaxRange, wallThickness)
        self.occupancyGridVisited, self.occupancyGridTotal = self.geom.new_counts()
        g = self.geom
        self.unitGridSize, self.lidarFOV, self.lidarMaxRange = g.unit, g.fov, g.maxRange
        self.numSamplesPerRev, self.angularStep = g.K, g.angularStep
        self.mapXLim, self.mapYLim = g.mapXLim, g.mapYLim

    def convertRealXYToMapIdx(self, x, y):
        return self.geom.map_index(x, y)

    def updateOccupancyGrid(self, reading, dTheta=0):
        update_counts(self.geom, self.occupancyGridVisited, self.occupancyGridTotal,
                      reading['x'], reading['y'], reading['theta'] + dTheta, reading['range'])


class ScanMatcher:
    """Call surface of Utils/ScanMatcher_OGBased.py:8."""

    def __init__(self, og, *params):
        self.og = og
        self.mp = MatcherParams(*params)
        self.trace = None

    def matchScan(self, reading, estMovingDist, estMovingTheta, count, matchMax=True, uniform=None):
        if count == 1:
            return reading, 1
        x, y, th, conf, cidx, fidx = match_scan(
            self.og.geom, self.mp, self.og.occupancyGridVisited, self.og.occupancyGridTotal,
            reading['x'], reading['y'], reading['theta'], reading['range'], estMovingDist, estMovingTheta,
            matchMax=matchMax, uniform=uniform, trace=self.trace)
        self.lastIdx = (cidx, fidx)
        return {'x': x, 'y': y, 'theta': th, 'range': reading['range']}, conf


class Particle:
    """Algorithm/FastSlam.py:64-140 (plotting omitted)."""

    def __init__(self, ogParameters, smParameters, geometry=None):
        mapX, mapY, initXY, unit, fov, maxRange, K, wall = ogParameters
        self.og = OccupancyGrid(mapX, mapY, initXY, unit, fov, K, maxRange, wall, geometry=geometry)
        self.sm = ScanMatcher(self.og, *smParameters)
        self.xTrajectory, self.yTrajectory = [], []
        self.weight = 1

    def update(self, reading, count, uniform=None):
        if count == 1:
            self.prevRawMovingTheta, self.prevMatchedMovingTheta = None, None
            matched, confidence = reading, 1
        else:
            ex, ey, eth, dist, estMovTh, rawMovTh = propose_pose(
                reading, self.prevMatchedReading, self.prevRawReading, self.prevRawMovingTheta,
                self.prevMatchedMovingTheta)
            est = {'x': ex, 'y': ey, 'theta': eth, 'range': reading['range']}
            matched, confidence = self.sm.matchScan(est, dist, estMovTh, count, matchMax=False, uniform=uniform)
            self.prevRawMovingTheta = rawMovTh
            self.prevMatchedMovingTheta = moving_heading(matched['x'], matched['y'],
                                                         self.xTrajectory[-1], self.yTrajectory[-1])
        self.xTrajectory.append(matched['x'])
        self.yTrajectory.append(matched['y'])
        self.og.updateOccupancyGrid(matched)
        self.prevMatchedReading, self.prevRawReading = matched, reading
        self.weight *= confidence


class ParticleFilter:
    """Algorithm/FastSlam.py:10-62.  Sector tables are shared between particles."""

    def __init__(self, numParticles, ogParameters, smParameters):
        import copy
        self._copy = copy
        self.numParticles = numParticles
        mapX, mapY, initXY, unit, fov, maxRange, K, wall = ogParameters
        geom = GridGeometry(mapX, mapY, initXY, unit, fov, K, maxRange, wall)
        self.particles = [Particle(ogParameters, smParameters, geometry=geom) for _ in range(numParticles)]

    def updateParticles(self, reading, count):
        for p in self.particles:
            p.update(reading, count)

    def normalizeWeights(self):
        for p, w in zip(self.particles, normalize([p.weight for p in self.particles])):
            p.weight = w

    def weightUnbalanced(self):
        self.normalizeWeights()
        flag, self.lastVariance = unbalanced([p.weight for p in self.particles])
        return flag

    def resample(self, uniforms=None):
        idx = resample_indices([p.weight for p in self.particles], uniforms)
        old = self.particles
        new = []
        for i in idx:
            geom = old[i].og.geom
            old[i].og.geom = None              # share the immutable tables instead of copying them
            q = self._copy.deepcopy(old[i])
            old[i].og.geom = geom
            q.og.geom = geom
            q.weight = 1 / self.numParticles
            new.append(q)
        self.particles = new
        self.lastResampleIdx = np.asarray(idx)
