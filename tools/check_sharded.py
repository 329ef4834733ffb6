"""torchrun --nproc-per-node 2 tools/check_sharded.py : sharding over ranks must be invisible in the results
(distributed.sharding_self_check: sharded vs unsharded filter, bit for bit, incl. a forced cross-rank resample)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
from slam_2d_lidar_scan_b200.distributed import sharding_self_check  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = sharding_self_check(dev)
if rank == 0:
    print("SHARDED_CHECK", "PASS" if ok else "FAIL", "world", world)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
