"""B200-native scan-match / FastSLAM hot path with the reference's Python class surface.

    from slam_2d_lidar_scan_b200 import OccupancyGrid, ScanMatcher, ParticleFilter, FastSLAM

(the directory name carries a hyphen; ``slam_2d_lidar_scan_b200.py`` at the repo root aliases it to an
importable name).  Every class keeps the constructor / method signatures of the reference
(Utils/OccupancyGrid.py, Utils/ScanMatcher_OGBased.py, Algorithm/FastSlam.py); the work runs in the sm_100a
kernels of ``csrc/`` through the C ABI declared in ``include/slam2d_b200.h``.  No CPU fallback.
"""
from . import _native
from .geometry import LidarGeometry
from .engine import MatcherEngine, gaussian_taps
from .grid import OccupancyGrid
from .matcher import ScanMatcher, updateEstimatedPose, updateTrajectory, getMovingTheta, readJson
from .fastslam import Particle, ParticleFilter, FastSLAM

__all__ = ["OccupancyGrid", "ScanMatcher", "Particle", "ParticleFilter", "FastSLAM", "LidarGeometry",
           "MatcherEngine", "updateEstimatedPose", "updateTrajectory", "getMovingTheta", "readJson"]
