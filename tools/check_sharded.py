"""torchrun --nproc-per-node 2 tools/check_sharded.py : sharding over ranks must be invisible in the results.

Every rank runs the sharded filter (N particles over WORLD_SIZE GPUs, one all-gather per step) and then, on its own
GPU, the unsharded filter with the same seed; poses, weights and the resample decision must agree bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
import slam_2d_lidar_scan_b200 as S  # noqa: E402
from slam_2d_lidar_scan_b200 import synthetic  # noqa: E402
from slam_2d_lidar_scan_b200.distributed import ShardedParticleFilter  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
spec = synthetic.config("c2")
N, steps = 8 * world, 9
scene = synthetic.make_scene(seed=1, steps=steps, K=spec["K"], fov=spec["og"][4], unit=spec["og"][3])


def run(cls, **kw):
    np.random.seed(77)
    pf = cls(N, spec["og"], spec["sm"], device=dev, **kw)
    local_pf = pf.local if hasattr(pf, "local") else pf
    og = S.OccupancyGrid(*local_pf.geom.args, _geometry=local_pf.geom)
    for fr in scene["warm"]:
        og.updateOccupancyGrid(fr)
    local_pf.grids.copy_(og.device_grid.unsqueeze(0).expand_as(local_pf.grids))
    log = []
    for count, fr in enumerate(scene["frames"][:steps], start=1):
        pf.updateParticles(fr, count)
        fired = pf.weightUnbalanced()
        if isinstance(pf, ShardedParticleFilter):
            w = pf._w.cpu().numpy().copy()
            poses = pf.poses().copy()
        else:
            w = pf.weights.cpu().numpy().copy()
            poses = pf.poses().copy()
        log.append((fired, poses, w))
        if count == 6:                       # force a resample to exercise the cross-rank lattice moves
            pf.resample()
    return pf, log


spf, a = run(ShardedParticleFilter)
ref, b = run(S.ParticleFilter)
ok = True
for (fa, pa, wa), (fb, pb, wb) in zip(a, b):
    ok &= fa == fb and np.array_equal(pa, pb) and np.array_equal(wa, wb)
# lattices after the forced resample + 3 more steps: local slice equals the unsharded filter's slice
lo, hi = spf.lo, spf.hi
ok &= bool(torch.equal(spf.local.grids, ref.grids[lo:hi]))
ok &= np.array_equal(spf.lastResampleIdx, ref.lastResampleIdx)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", "world", world, "particles", N,
          "distinct poses", len(np.unique(a[-1][1][:, 0])))
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
