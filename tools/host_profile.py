"""cProfile of the per-step host path (python tools/host_profile.py [workload] [particles] [steps]) -- what the host does
between the trigger read-back of one step and the match launch of the next is GPU idle time in the end-to-end number."""
import cProfile
import os
import pstats
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
import slam_2d_lidar_scan_b200 as S  # noqa: E402
from slam_2d_lidar_scan_b200 import synthetic  # noqa: E402
from slam_2d_lidar_scan_b200.distributed import ShardedParticleFilter  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c3"
spec = synthetic.config(workload)
n = int(sys.argv[2]) if len(sys.argv) > 2 else spec["N"]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 60
scene = synthetic.make_scene(seed=0, steps=steps + 12, K=spec["K"], fov=spec["og"][4], unit=spec["og"][3], stride=0.35)
np.random.seed(0)
spf = ShardedParticleFilter(n, spec["og"], spec["sm"])
pf = spf.local
pf.keepTrajectory = False
pf.expandMaps = False
pf.ignoreMissingHeading = True
og = S.OccupancyGrid(*pf.geom.args, _geometry=pf.geom)
for fr in scene["warm"]:
    og.updateOccupancyGrid(fr)
pf.grids.copy_(og.device_grid.unsqueeze(0).expand_as(pf.grids))
frames = scene["frames"]
for count in range(1, 9):
    spf.updateParticles(frames[count - 1], count)
    spf.weightUnbalanced()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for count in range(9, 9 + steps):
    spf.updateParticles(frames[count - 1], count)
    spf.weightUnbalanced()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
