"""Host-side lattice and lidar-sector tables, built with the reference's own float64 expressions and uploaded
once (shared by every particle; the reference rebuilds them per particle).

Reference: Utils/OccupancyGrid.py:7-45 (OccupancyGrid.__init__, spokesGrid).  The per-spoke cell lists of
itemizeSpokesGrid (:47-57) are not materialised: the update kernel inverts the sector lookup instead.
"""
import numpy as np
import torch

from . import _native as nat


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("slam-2d-lidar-scan_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())


class LidarGeometry:
    def __init__(self, mapXLength, mapYLength, initXY, unitGridSize, lidarFOV, numSamplesPerRev, lidarMaxRange,
                 wallThickness, device=None, _cells=None):
        xNum = int(mapXLength / unitGridSize)
        yNum = int(mapYLength / unitGridSize)
        if _cells is not None:          # grown(): the cell count is given exactly (no float division)
            xNum = yNum = int(_cells)
        if xNum != yNum:
            # the reference mixes xNum / yNum (OccupancyGrid.py:11,13); only square maps are well defined
            raise NotImplementedError("only square maps are supported (mapXLength == mapYLength)")
        if numSamplesPerRev > nat.MAX_BEAMS:
            raise NotImplementedError("at most %d beams per scan" % nat.MAX_BEAMS)
        self.args = (mapXLength, mapYLength, dict(initXY), unitGridSize, lidarFOV, numSamplesPerRev, lidarMaxRange,
                     wallThickness)
        self.initXY = dict(initXY)
        self.cells = xNum
        self.unitGridSize = unitGridSize
        self.lidarFOV = lidarFOV
        self.lidarMaxRange = lidarMaxRange
        self.wallThickness = wallThickness
        self.numSamplesPerRev = numSamplesPerRev
        half = xNum * unitGridSize / 2
        self.gridX = np.linspace(-half, half, num=xNum + 1) + initXY['x']          # OccupancyGrid.py:10
        self.gridY = np.linspace(-half, half, num=yNum + 1) + initXY['y']          # :11
        self.G = xNum + 1
        self.pitch = (self.G + 3) // 4 * 4
        self.mapXLim = [self.gridX[0], self.gridX[-1]]                             # :19-20
        self.mapYLim = [self.gridY[0], self.gridY[-1]]
        self.angularStep = lidarFOV / numSamplesPerRev                             # :22
        self.numSpokes = int(np.rint(2 * np.pi / self.angularStep))                # :23
        self.spokesStartIdx = int(((self.numSpokes / 2 - numSamplesPerRev) / 2) % self.numSpokes)   # :30
        self._build_sector_tables()
        self.device = require_cuda(device)
        self._updWs, self._updWsN = {}, {}
        dev = self.device
        self.d_gridX = torch.from_numpy(self.gridX).to(dev)
        self.d_gridY = torch.from_numpy(self.gridY).to(dev)
        self.d_sector = torch.from_numpy(self.sector.astype(np.int16)).to(dev).contiguous()
        self.d_radius = torch.from_numpy(self.radius).to(dev).contiguous()
        self.d_localAxis = torch.from_numpy(self.localAxis).to(dev)
        g = nat.Geometry()
        g.G, g.pitch, g.K, g.L = self.G, self.pitch, numSamplesPerRev, self.L
        g.numSpokes, g.spokesStartIdx = self.numSpokes, self.spokesStartIdx
        g.unit = unitGridSize
        g.mapX0, g.mapX1 = self.mapXLim
        g.mapY0, g.mapY1 = self.mapYLim
        g.fovHalf = lidarFOV / 2
        g.maxRange = lidarMaxRange
        g.wallHalf = wallThickness / 2
        g.d_gridX, g.d_gridY = self.d_gridX.data_ptr(), self.d_gridY.data_ptr()
        g.d_sector, g.d_radius = self.d_sector.data_ptr(), self.d_radius.data_ptr()
        g.d_localAxis = self.d_localAxis.data_ptr()
        self.c = g

    def _build_sector_tables(self):
        """Bearing sector (0 = -y, counter-clockwise) and radius of each lidar-local cell (:32-45)."""
        n = int(self.lidarMaxRange / self.unitGridSize)
        L = 2 * n + 1
        axis = np.linspace(-self.lidarMaxRange, self.lidarMaxRange, L)
        xg, yg = np.meshgrid(axis, axis)
        sec = np.zeros((L, L))
        with np.errstate(divide='ignore', invalid='ignore'):
            east = np.rint((np.pi / 2 + np.arctan(yg[:, n + 1:] / xg[:, n + 1:])) / np.pi / 2 * self.numSpokes - 0.5)
        sec[:, n + 1:] = east.astype(int)
        sec[:, :n] = sec[::-1, ::-1][:, :n] + int(self.numSpokes / 2)     # west = east rotated by half a turn
        sec[n + 1:, n] = int(self.numSpokes / 2)                          # +y half of the centre column
        self.L = L
        self.localAxis = axis
        self.sector = sec.astype(np.int32)
        self.radius = np.sqrt(xg ** 2 + yg ** 2)

    def grown(self):
        """The lattice of a map twice as long around the same centre: exactly the lattice a map pre-sized to that
        length would have (same linspace expression), so an expanded map IS a pre-sized one.  -> (geometry, offset)
        where old cell (i, j) is new cell (i + offset, j + offset).  Map expansion, OccupancyGrid.py:59-125."""
        if self.cells % 2:
            raise NotImplementedError("map expansion needs an even number of cells per side")
        new = LidarGeometry(2 * self.args[0], 2 * self.args[1], self.initXY, self.unitGridSize, self.lidarFOV,
                            self.numSamplesPerRev, self.lidarMaxRange, self.wallThickness, device=self.device,
                            _cells=2 * self.cells)
        return new, self.cells // 2

    def rehome(self, grids, new, offset):
        """Counts of lattices ``grids`` [n][G][pitch][2] copied into fresh lattices of geometry ``new`` (device copy)."""
        out = new.new_grids(grids.shape[0])
        out[:, offset:offset + self.G, offset:offset + self.G, :] = grids[:, :, :self.G, :]
        return out

    def contains(self, x0, x1, y0, y1):
        return x0 >= self.mapXLim[0] and x1 <= self.mapXLim[1] and y0 >= self.mapYLim[0] and y1 <= self.mapYLim[1]

    def update_workspace(self, n, stream=None):
        """Device scratch of slam_update_grid for n particles, one buffer per CUDA stream (grown on demand)."""
        key = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        ws = self._updWs.get(key)
        if ws is not None and self._updWsN.get(key, 0) >= n:
            return ws
        need = nat.lib.slam_update_workspace_bytes(n)
        self._updWsN[key] = n
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._updWs[key] = ws
        return ws

    def mapIndex(self, x, y):
        """convertRealXYToMapIdx (OccupancyGrid.py:102-106)."""
        xi = np.rint((np.asarray(x) - self.mapXLim[0]) / self.unitGridSize).astype(int)
        yi = np.rint((np.asarray(y) - self.mapYLim[0]) / self.unitGridSize).astype(int)
        return xi, yi

    def new_grids(self, n):
        """[n][G][pitch][2] float32 lattices initialised to (visited, total) = (1, 2) (:13-14)."""
        g = torch.empty((n, self.G, self.pitch, 2), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(nat.lib.slam_grid_init(self.c, g.data_ptr(), n, torch.cuda.current_stream(self.device).cuda_stream))
        return g
