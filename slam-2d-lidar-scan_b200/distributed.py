"""Particles sharded over GPUs, one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

Particles are independent units: each rank owns a contiguous block of N/P particles and
their lattices; matchScan + map update need no communication.  Per step there is ONE collective -- an all-gather
of [N_local][4] float64 (unnormalised weight, x, y, theta) -- after which normalisation, the resample trigger and
the resample indices are computed redundantly and identically on every rank (same sequential float64 order as
FastSlam.py:30-62).  Only when a resample fires do lattices move between ranks (point-to-point, source -> slot).

Every rank seeds numpy identically and draws the uniforms of ALL N particles, using its own slice, so the result
is bit-identical to the single-GPU run of the same N particles (partitioning is invisible).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _native as nat
from .engine import StepResult, raise_for_status, _stream
from .fastslam import ParticleFilter


def shard_bounds(numParticles, world):
    if numParticles % world:
        raise ValueError("numParticles must be divisible by the number of ranks")
    n = numParticles // world
    return [(r * n, (r + 1) * n) for r in range(world)]


def plan_resample_transfers(idx, nLocal, world):
    """For global resample indices idx[i] (destination slot i <- source idx[i]) return, per rank, the lists
    local[(dstLocal, srcLocal)], sends[(dstRank, srcLocal, dstGlobal)], recvs[(srcRank, dstLocal, dstGlobal)],
    ordered by destination slot so that matching sends/recvs pair up deterministically."""
    plans = [dict(local=[], sends=[], recvs=[]) for _ in range(world)]
    for i, s in enumerate(int(v) for v in idx):
        d, sr = i // nLocal, s // nLocal
        if d == sr:
            plans[d]["local"].append((i % nLocal, s % nLocal))
        else:
            plans[sr]["sends"].append((d, s % nLocal, i))
            plans[d]["recvs"].append((sr, i % nLocal, i))
    return plans


class ShardedParticleFilter:
    """ParticleFilter surface (updateParticles / weightUnbalanced / resample) over ranks of a process group."""

    def __init__(self, numParticles, ogParameters, smParameters, *, device=None, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.numParticles = numParticles
        self.lo, self.hi = shard_bounds(numParticles, self.world)[self.rank]
        self.local = ParticleFilter(self.hi - self.lo, ogParameters, smParameters, device=device)
        dev = self.local.geom.device
        nL = self.hi - self.lo
        # all-gather payload per rank: nL rows (unnormalised weight, x, y, theta) + one row carrying the OR of the
        # rank's status words, so that every rank raises the same exception in the same step (no hung collectives)
        self._mine = torch.zeros((nL + 1, 4), dtype=torch.float64, device=dev)
        self._all = torch.zeros((self.world, nL + 1, 4), dtype=torch.float64, device=dev)
        self._stRanks = torch.zeros(self.world, dtype=torch.int32, device=dev)
        self._w = torch.zeros(numParticles, dtype=torch.float64, device=dev)
        self._res = StepResult(dev)
        self._out = self._res.out
        self._cdf = torch.zeros(numParticles, dtype=torch.float64, device=dev)
        self._ridx = torch.zeros(numParticles, dtype=torch.int32, device=dev)
        self.lastVariance = None
        self.lastResampleIdx = None

    def updateParticles(self, reading, count):
        u = np.random.random_sample(self.numParticles) if count > 1 else None     # the global stream, all ranks
        self.local._update(0, self.hi - self.lo, reading, count, uniforms=None if u is None else u[self.lo:self.hi])

    def gather_and_normalize(self):
        """The step's single collective + the replicated sequential normalisation.  No host synchronisation."""
        pf = self.local
        nL, dev = self.hi - self.lo, pf.geom.device
        pf._res.reduce_status(pf.status, dev)
        self._mine[:nL, 0] = pf.weights
        self._mine[:nL, 1:] = pf.prevMatched
        self._mine[nL, 0:1] = pf._res.bits.to(torch.float64)
        if self.world > 1:
            dist.all_gather_into_tensor(self._all.view(-1, 4), self._mine, group=self.group)
        else:
            self._all[0].copy_(self._mine)
        self._w.copy_(self._all[:, :nL, 0].reshape(-1))
        self._stRanks.copy_(self._all[:, nL, 0])
        with torch.cuda.device(dev):
            nat.check(nat.lib.slam_normalize_weights(self.numParticles, self._w.data_ptr(), self._out.data_ptr(),
                                                     _stream(dev)))
        self._res.reduce_status(self._stRanks, dev)
        pf.kernelLaunches += 3
        pf.weights.copy_(self._w[self.lo:self.hi])

    def weightUnbalanced(self):
        self.gather_and_normalize()
        var, fired, bits = self._res.fetch()
        self.local.d2hBytes += 24
        raise_for_status(bits)
        self.lastVariance = var
        return fired

    def poses(self):
        """[N][3] poses of all particles as of the last gather (host numpy)."""
        return self._all[:, :self.hi - self.lo, 1:].reshape(-1, 3).cpu().numpy()

    def resample(self):
        """Global multinomial resample (FastSlam.py:50-62); lattices whose source lives on another rank move P2P."""
        pf, n, nL = self.local, self.numParticles, self.hi - self.lo
        dev = pf.geom.device
        u = torch.from_numpy(np.random.random_sample(n)).to(dev)                  # same draw on every rank
        nat.check(nat.lib.slam_resample_indices(n, self._w.data_ptr(), u.data_ptr(), self._cdf.data_ptr(),
                                                self._ridx.data_ptr(), _stream(dev)))
        idx = self._ridx.cpu().numpy()
        plan = plan_resample_transfers(idx, nL, self.world)[self.rank]
        state = torch.cat([pf.prevMatched, pf.prevHeading.view(-1, 1), pf.hasHeading.to(torch.float64).view(-1, 1)], 1)
        newGrids, newState = torch.empty_like(pf.grids), torch.empty_like(state)
        for d, s in plan["local"]:
            newGrids[d].copy_(pf.grids[s])
            newState[d].copy_(state[s])
        ops = []
        for dstRank, s, tag in plan["sends"]:
            ops.append(dist.P2POp(dist.isend, pf.grids[s], dstRank, group=self.group))
            ops.append(dist.P2POp(dist.isend, state[s], dstRank, group=self.group))
        for srcRank, d, tag in plan["recvs"]:
            ops.append(dist.P2POp(dist.irecv, newGrids[d], srcRank, group=self.group))
            ops.append(dist.P2POp(dist.irecv, newState[d], srcRank, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        pf.grids.copy_(newGrids)
        del newGrids
        pf.prevMatched.copy_(newState[:, :3])
        pf.prevHeading.copy_(newState[:, 3])
        pf.hasHeading.copy_(newState[:, 4].to(torch.int32))
        pf.weights.fill_(1.0 / n)
        src = [int(idx[i]) for i in range(self.lo, self.hi)]
        # trajectories / raw-odometry records are identical host data on every rank except per-particle history
        pf._traj = []           # per-particle history does not follow a particle across ranks (documented)
        self.lastResampleIdx = idx.astype(np.int64)
        return src
