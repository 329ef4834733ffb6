import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.build()
import slam_2d_lidar_scan_b200 as S
from slam_2d_lidar_scan_b200 import synthetic
spec = synthetic.config("c3")
scene = synthetic.make_scene(seed=0, steps=12, K=spec["K"], fov=spec["og"][4], unit=spec["og"][3], stride=0.35)
np.random.seed(1234)
pf = S.ParticleFilter(1024, spec["og"], spec["sm"])
og = S.OccupancyGrid(*pf.geom.args, _geometry=pf.geom)
for fr in scene["warm"]:
    og.updateOccupancyGrid(fr)
pf.load_grid(og.device_grid)
for count, fr in enumerate(scene["frames"][:12], start=1):
    before = pf.prevMatched.clone()
    pf._update(0, 1024, fr, count)
    torch.cuda.synchronize()
    mv = (pf.prevMatched[:, :2] - before[:, :2]).norm(dim=1)
    print(count, "raw", fr["x"], fr["y"], "hasHeading==0:", int((pf.hasHeading == 0).sum()), "zero moves:", int((mv == 0).sum()),
          "status", int(pf.status.max()), "idx sample", pf._idx[:3].cpu().tolist(), "move min/mean", float(mv.min()), float(mv.mean()))
