"""Run a few FastSLAM steps of a workload (for ncu): python tools/run_steps.py [workload] [particles] [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
import slam_2d_lidar_scan_b200 as S  # noqa: E402
from slam_2d_lidar_scan_b200 import synthetic  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c3"
spec = synthetic.config(workload)
n = int(sys.argv[2]) if len(sys.argv) > 2 else spec["N"]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
scene = synthetic.make_scene(seed=0, steps=steps + 2, K=spec["K"], fov=spec["og"][4], unit=spec["og"][3])
np.random.seed(0)
pf = S.ParticleFilter(n, spec["og"], spec["sm"])
pf.keepTrajectory = False
og = S.OccupancyGrid(*pf.geom.args, _geometry=pf.geom)
for fr in scene["warm"]:
    og.updateOccupancyGrid(fr)
pf.grids.copy_(og.device_grid.unsqueeze(0).expand_as(pf.grids))
for count, fr in enumerate(scene["frames"][:steps], start=1):
    pf.updateParticles(fr, count)
    pf.weightUnbalanced()
torch.cuda.synchronize()
print("ok", pf.poses()[:2])
