"""Per-phase SM-cycle breakdown of the fused match kernel (clock64 at phase boundaries, thread 0 of each CTA).
Usage (GPU box): python tools/phase_cycles.py [workload] [particles]"""
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
after = int(sys.argv[3]) if len(sys.argv) > 3 else 4        # steps before the two measured ones (maps fill up with time)
stride = float(sys.argv[4]) if len(sys.argv) > 4 else 0.25
scene = synthetic.make_scene(seed=0, steps=after + 4, K=spec["K"], fov=spec["og"][4], unit=spec["og"][3], stride=stride)
np.random.seed(0)
pf = S.ParticleFilter(n, spec["og"], spec["sm"])
pf.keepTrajectory = False
og = S.OccupancyGrid(*pf.geom.args, _geometry=pf.geom)
for fr in scene["warm"]:
    og.updateOccupancyGrid(fr)
pf.grids.copy_(og.device_grid.unsqueeze(0).expand_as(pf.grids))
nat = S._native
ctas = nat.lib.slam_matcher_num_ctas(pf.engine.handle)
cyc = torch.zeros((ctas, 48), dtype=torch.int64, device=pf.geom.device)
plan = (nat.C.c_int * 8)()
for s in range(2):
    nat.lib.slam_matcher_plan(pf.engine.handle, s, nat.C.byref(plan))
    print("stage", s, "PInSmem,scoresInSmem,bitsInSmem,R,TB,Ppitch,smemBytes,slotKB =", list(plan))
pf.ignoreMissingHeading = True
for count, fr in enumerate(scene["frames"][:after + 2], start=1):
    if count == after + 1:
        nat.lib.slam_matcher_set_debug(pf.engine.handle, cyc.data_ptr(), 0)
        cyc.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pf.matchEvents = []
    pf.updateParticles(fr, count)
    pf.weightUnbalanced()
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in pf.matchEvents]
c = cyc.cpu().numpy().astype(np.float64)
names = ["window", "blur+clamp", "points", "lists(sort)", "correlate", "select", "wait-union", "-"]
per = c.sum(0) / (2 * n)           # two timed launches
tot = per.sum()
print("match kernel ms per launch:", ms)
for st in range(2):
    for k in range(7 if st == 0 else 6):
        v = per[8 * st + k]
        print("%-7s %-12s %10.0f cycles/particle  %5.1f %%" % ("coarse" if st == 0 else "fine", names[k], v, 100 * v / tot))
sub = ["B clear+maps", "C scatter", "C transpose", "-", "D1 dilate", "D2 blur (own tiles)", "-", "-", "H argmax", "H exp", "H sum", "H cdf"]
for st in range(2):
    print("  sub-phases %s: " % ("coarse" if st == 0 else "fine") + ", ".join("%s %.0f" % (sub[k], per[16 + 16 * st + k]) for k in range(12) if sub[k] != "-"))
print("cold start per CTA: compute-side pack %.0f, then wait for the stream warps %.0f cycles" % (c[:, 29].sum() / (2 * ctas), c[:, 28].sum() / (2 * ctas)))
print("fine branch-and-bound: %.1f %% of the points of the evaluated hypotheses were gathered" % (100.0 * c[:, 44].sum() / max(c[:, 45].sum(), 1)))
print("fine field: %.0f active tiles, %.0f active cells per particle" % (c[:, 46].sum() / (2 * n), c[:, 47].sum() / (2 * n)))
per[16:] = 0
print("stream warp: TMA wait %.0f, pack %.0f, bitmap-free wait %.0f cycles/particle" % (per[7], per[14], per[15]))
per[7] = per[14] = per[15] = 0
tot = per.sum()
print("total %.0f cycles/particle; status max %d" % (tot, int(pf.status.max().item())))
