"""OccupancyGrid with the reference's constructor / method surface (Utils/OccupancyGrid.py:6-175), backed by one
device-resident lattice (or a view of one slot of a ParticleFilter's batch)."""
import numpy as np
import torch

from . import _native as nat
from .engine import raise_for_status, update_grids
from .geometry import LidarGeometry


def scan_reach(ranges, maxRange, wallThickness):
    """Largest distance from the pose at which this scan can touch a cell (OccupancyGrid.py:138-143): a beam empties
    r < range - wall/2 (only if range < maxRange) and marks range - wall/2 < r < range + wall/2; the lidar-local patch
    ends at sqrt(2) * maxRange, so longer beams (the logs' 81.83 m sentinel) touch nothing."""
    r = np.asarray(ranges, dtype=np.float64)
    hit = r[r - wallThickness / 2 < 1.4143 * maxRange]
    reach = float(hit.max() + wallThickness / 2) if hit.size else 0.0
    return min(reach, 1.4143 * maxRange)


class OccupancyGrid:
    def __init__(self, mapXLength, mapYLength, initXY, unitGridSize, lidarFOV, numSamplesPerRev, lidarMaxRange,
                 wallThickness, *, device=None, _geometry=None, _grids=None, _slot=0, _slotmap=None):
        self.geom = _geometry or LidarGeometry(mapXLength, mapYLength, initXY, unitGridSize, lidarFOV,
                                               numSamplesPerRev, lidarMaxRange, wallThickness, device=device)
        g = self.geom
        self._grids = g.new_grids(1) if _grids is None else _grids      # [N][G][pitch][2] float32
        self._slot = _slot
        self._slotmap = _slotmap          # particle -> lattice table of the owning filter (copy-elided resampling)
        self.unitGridSize = g.unitGridSize
        self.lidarFOV = g.lidarFOV
        self.lidarMaxRange = g.lidarMaxRange
        self.wallThickness = g.wallThickness
        self.numSamplesPerRev = g.numSamplesPerRev
        self.angularStep = g.angularStep
        self.numSpokes = g.numSpokes
        self.spokesStartIdx = g.spokesStartIdx
        self.mapXLim = list(g.mapXLim)
        self.mapYLim = list(g.mapYLim)
        dev = g.device
        self._status = torch.zeros(1, dtype=torch.int32, device=dev)
        self._pose = torch.zeros(3, dtype=torch.float64, device=dev)
        self._ranges = torch.zeros(g.numSamplesPerRev, dtype=torch.float64, device=dev)

    # ---- device views
    @property
    def device_grid(self):
        """[G][pitch][2] float32 view of this map's (visited, total) counts."""
        return self._grids[self._slot if self._slotmap is None else int(self._slotmap()[self._slot])]

    def _counts(self, ch):
        return self.device_grid[:, :self.geom.G, ch].to(torch.float64).cpu().numpy()

    @property
    def occupancyGridVisited(self):
        return self._counts(0)

    @property
    def occupancyGridTotal(self):
        return self._counts(1)

    @property
    def OccupancyGridX(self):
        return np.meshgrid(self.geom.gridX, self.geom.gridY)[0]

    @property
    def OccupancyGridY(self):
        return np.meshgrid(self.geom.gridX, self.geom.gridY)[1]

    # ---- reference methods
    def convertRealXYToMapIdx(self, x, y):
        return self.geom.mapIndex(x, y)

    def checkMapToExpand(self, x, y):
        x, y = np.asarray(x), np.asarray(y)
        if np.any(x < self.mapXLim[0]):
            return 1
        if np.any(x > self.mapXLim[1]):
            return 2
        if np.any(y < self.mapYLim[0]):
            return 3
        if np.any(y > self.mapYLim[1]):
            return 4
        return -1

    def checkAndExapndOG(self, x, y):
        """Grow the map until it holds the points (OccupancyGrid.py:120-125).  Deviation from the reference, whose
        expansion has defects (non-uniform inserted coordinates, stale indices): the map doubles around
        its centre and becomes exactly the lattice a map pre-sized to that length has ("virtually pre-sized")."""
        while self.checkMapToExpand(x, y) != -1:
            self.expandOccupancyGrid(self.checkMapToExpand(x, y))

    def expandOccupancyGrid(self, expandDirection):
        """OccupancyGrid.py:93-101.  Every direction grows the map symmetrically (see checkAndExapndOG)."""
        if self._slotmap is not None:
            raise IndexError("a particle's map grows with its filter (ParticleFilter expands all lattices together)")
        new, off = self.geom.grown()
        self._grids = self.geom.rehome(self._grids, new, off)
        self._adopt(new)

    def _adopt(self, geom):
        self.geom = geom
        self.mapXLim = list(geom.mapXLim)
        self.mapYLim = list(geom.mapYLim)

    def _touched_extent(self, reading, dTheta=0):
        """World bounding box of the cells this scan empties or marks (OccupancyGrid.py:131-145), or None."""
        g = self.geom
        ranges = np.asarray(reading['range'], dtype=np.float64)
        off = int(np.rint((reading['theta'] + dTheta) / (2 * np.pi) * g.numSpokes))              # :131
        beam = (g.sector - (g.spokesStartIdx + off)) % g.numSpokes                              # inverse of :134
        valid = beam < g.numSamplesPerRev
        b = np.where(valid, beam, 0)
        lo, hi = (ranges - g.wallThickness / 2)[b], (ranges + g.wallThickness / 2)[b]
        touched = valid & (((ranges[b] < g.lidarMaxRange) & (g.radius < lo)) | ((g.radius > lo) & (g.radius < hi)))
        ys, xs = np.nonzero(touched)
        if ys.size == 0:
            return None
        ax = g.localAxis
        return (reading['x'] + ax[xs.min()], reading['x'] + ax[xs.max()], reading['y'] + ax[ys.min()],
                reading['y'] + ax[ys.max()])

    def _cover(self, x0, x1, y0, y1):
        """Expand (standalone maps only) until the box lies inside the map."""
        if not self.geom.contains(x0, x1, y0, y1):
            self.checkAndExapndOG(np.array([x0, x1]), np.array([y0, y1]))

    def updateOccupancyGrid(self, reading, dTheta=0, update=True):
        """OccupancyGrid.py:127-152 on the GPU (slam_update_grid)."""
        if not update:
            raise NotImplementedError("update=False (coordinate lists) has no caller in the reference")
        if self._slotmap is None:
            m = scan_reach(reading['range'], self.lidarMaxRange, self.wallThickness) + self.unitGridSize
            x, y = reading['x'], reading['y']
            if not self.geom.contains(x - m, x + m, y - m, y + m):
                # close to the border: decide with the cells the scan really touches (a map that is large enough is
                # never grown -- growing changes the lattice's coordinate rounding)
                box = self._touched_extent(reading, dTheta)
                if box is not None:
                    self._cover(*box)
        pose = torch.tensor([reading['x'], reading['y'], reading['theta'] + dTheta], dtype=torch.float64)
        rng = torch.as_tensor(np.asarray(reading['range'], dtype=np.float64))
        self._pose.copy_(pose)
        self._ranges.copy_(rng)
        self._status.zero_()
        update_grids(self.geom, self.device_grid, 1, self._ranges, self._pose, self._status)
        raise_for_status(int(self._status.item()))

    def plotOccupancyGrid(self, xRange=None, yRange=None, plotThreshold=True):
        raise NotImplementedError("plotting is out of scope; read occupancyGridVisited / occupancyGridTotal")
