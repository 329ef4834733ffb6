"""OccupancyGrid with the reference's constructor / method surface (Utils/OccupancyGrid.py:6-175), backed by one
device-resident lattice (or a view of one slot of a ParticleFilter's batch)."""
import numpy as np
import torch

from . import _native as nat
from .engine import raise_for_status, update_grids
from .geometry import LidarGeometry


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
        """The reference grows the map here (OccupancyGrid.py:120-125); this implementation works on pre-sized
        lattices (SURVEY.md A.8 lists the reference's expansion defects), so leaving the map is an error."""
        if self.checkMapToExpand(x, y) != -1:
            raise IndexError("coordinates outside the pre-sized map; construct the OccupancyGrid large enough")

    def expandOccupancyGrid(self, expandDirection):
        raise NotImplementedError("map expansion is out of scope (pre-size the map)")

    def updateOccupancyGrid(self, reading, dTheta=0, update=True):
        """OccupancyGrid.py:127-152 on the GPU (slam_update_grid)."""
        if not update:
            raise NotImplementedError("update=False (coordinate lists) has no caller in the reference")
        pose = torch.tensor([reading['x'], reading['y'], reading['theta'] + dTheta], dtype=torch.float64)
        rng = torch.as_tensor(np.asarray(reading['range'], dtype=np.float64))
        self._pose.copy_(pose)
        self._ranges.copy_(rng)
        self._status.zero_()
        update_grids(self.geom, self.device_grid, 1, self._ranges, self._pose, self._status)
        raise_for_status(int(self._status.item()))

    def plotOccupancyGrid(self, xRange=None, yRange=None, plotThreshold=True):
        raise NotImplementedError("plotting is out of scope; read occupancyGridVisited / occupancyGridTotal")
