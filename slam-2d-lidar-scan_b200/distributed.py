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
from .engine import StepResult, SideTrigger, raise_for_status, copy_lattices, _stream
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
        self.local._tightenBound = False     # every rank must grow its maps in the same step (bound from the readings only)
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
        # The step's collective + replicated normalisation depend on the matched poses / weights only, not on the map
        # update that follows them: they run on the local filter's side stream next to the update kernels (the
        # single-thread sequential normalisation over all N_total weights and the all-gather latency leave the critical
        # path), see engine.SideTrigger.
        if self.world > 1:
            self.local._side = SideTrigger(dev, self._res)
            self.local._sideBody = self._gather_body
        else:                                 # one rank: nothing to gather -- the local filter's own trigger
            self._res, self._w = self.local._res, self.local._wn

    def updateParticles(self, reading, count):
        u = np.random.random_sample(self.numParticles) if count > 1 else None     # the global stream, all ranks
        self.local._update(0, self.hi - self.lo, reading, count, uniforms=None if u is None else u[self.lo:self.hi])

    def gather_and_normalize(self):
        """The step's single collective + the replicated sequential normalisation, without host synchronisation and
        without waiting on the device either: when the last launch started them on the side stream this only notes that
        the normalised weights must be copied back before the next step's weight update (ParticleFilter._settle), so the
        collective overlaps the map update AND the next step's match kernel.  flush() completes it on the caller's stream.
        Status bits raised by the map update itself (SLAM_ST_SCAN_OUTSIDE_MAP) are sticky and travel with the next gather."""
        if not self.local._side.pending:
            self.local._sideBody()
        self.local._copyBack = self._w[self.lo:self.hi]

    def flush(self):
        """Stream-side completion of a gather_and_normalize(): Particle.weight holds the normalised weights afterwards."""
        self.local._settle()

    def _gather_body(self):
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
        self._res.trigger(self._w, self._w, self._stRanks, dev)      # normalise in place + trigger + OR of the ranks' bits
        pf.kernelLaunches += 2

    def weightUnbalanced(self):
        if self.local._side.pending:
            var, fired, bits = self.local._side.fetch()
        else:
            self.local._sideBody()
            var, fired, bits = self._res.fetch()
        self.local._copyBack = None
        self.local.weights.copy_(self._w[self.lo:self.hi])       # Particle.weight is raw until here
        self.local.d2hBytes += 24
        if self.local.ignoreMissingHeading:
            bits &= ~nat.ST_HEADING_MISSING
        raise_for_status(bits & ~self.local.ignoreStatusBits)
        self.lastVariance = var
        return fired

    def poses(self):
        """[N][3] poses of all particles as of the last gather (host numpy)."""
        if self.world == 1:
            return self.local.poses()
        return self._all[:, :self.hi - self.lo, 1:].reshape(-1, 3).cpu().numpy()

    def resample(self):
        """Global multinomial resample (FastSlam.py:50-62).  Lattices whose source lives on another rank move
        point-to-point into staging buffers; locally, ownership of a chosen particle's lattice is handed over through
        the slot table and only the extra copies of multiply-chosen particles are physically copied
        (engine.plan_copy_elided).  Per-particle state and trajectory history travel with the particle."""
        pf, n, nL = self.local, self.numParticles, self.hi - self.lo
        dev = pf.geom.device
        u = torch.from_numpy(np.random.random_sample(n)).to(dev)                  # same draw on every rank
        with torch.cuda.device(dev):
            nat.check(nat.lib.slam_resample_indices(n, self._w.data_ptr(), u.data_ptr(), self._cdf.data_ptr(),
                                                    self._ridx.data_ptr(), _stream(dev)))
        idx = self._ridx.cpu().numpy()
        plan = plan_resample_transfers(idx, nL, self.world)[self.rank]
        hist = torch.stack(pf._traj, 0) if pf._traj else None                    # [T][nL][2] matched positions so far
        cols = [pf.prevMatched, pf.prevHeading.view(-1, 1), pf.hasHeading.to(torch.float64).view(-1, 1)]
        if hist is not None:
            cols.append(hist.permute(1, 0, 2).reshape(nL, -1))
        state = torch.cat(cols, 1).contiguous()
        newState = torch.empty_like(state)
        # remote sources: lattices arrive in staging buffers (their final lattice may still be a send source)
        staging = {d: torch.empty_like(pf.grids[0]) for _, d, _ in plan["recvs"]}
        ops = []
        for dstRank, sLoc, tag in plan["sends"]:
            ops.append(dist.P2POp(dist.isend, pf.lattice(sLoc), dstRank, group=self.group))
            ops.append(dist.P2POp(dist.isend, state[sLoc], dstRank, group=self.group))
        for srcRank, d, tag in plan["recvs"]:
            ops.append(dist.P2POp(dist.irecv, staging[d], srcRank, group=self.group))
            ops.append(dist.P2POp(dist.irecv, newState[d], srcRank, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        # local part: new local particle d <- old local particle s (ownership hand-over / elided copies); destinations
        # with a remote source take the lattices nobody (locally) chose
        local = dict(plan["local"])                                              # dstLocal -> srcLocal
        taken = np.zeros(nL, dtype=bool)
        newSlots = np.empty(nL, dtype=np.int32)
        extra, copies = [], []
        for d in range(nL):
            sLoc = local.get(d)
            if sLoc is not None and not taken[sLoc]:
                taken[sLoc] = True
                newSlots[d] = pf._slots_h[sLoc]
            else:
                extra.append(d)
            if sLoc is not None:
                newState[d].copy_(state[sLoc])
        free = list(pf._slots_h[~taken])
        for d, f in zip(extra, free):
            newSlots[d] = f
            if d in staging:
                pf.grids[int(f)].copy_(staging[d])
            else:
                copies.append((int(pf._slots_h[local[d]]), int(f)))
        copy_lattices(pf.geom, pf.grids, copies)
        pf._slots_h = newSlots
        pf.slots.copy_(torch.from_numpy(newSlots))
        pf.resampleCopies += len(copies) + len(staging)
        pf.lastResampleCopies = len(copies) + len(staging)
        pf.prevMatched = newState[:, :3].contiguous()
        pf.prevHeading = newState[:, 3].contiguous()
        pf.hasHeading = newState[:, 4].to(torch.int32).contiguous()
        if hist is not None:
            pf._traj = list(newState[:, 5:].reshape(nL, hist.shape[0], 2).permute(1, 0, 2).contiguous().unbind(0))
        pf.weights.fill_(1.0 / n)
        src = [int(idx[i]) for i in range(self.lo, self.hi)]
        self.lastResampleIdx = idx.astype(np.int64)
        return src


def sharding_self_check(dev, steps=9, perRank=8, workload="c2", seed=77):
    """Sharding must be invisible in the results: every rank runs the sharded filter (perRank particles per rank, one
    all-gather per step, one forced cross-rank resample) and then the unsharded filter with the same seed on its own
    GPU; poses, weights, trigger decisions, resample indices and the rank's lattices must agree bit for bit on every
    rank.  Collective (call on all ranks); returns the global verdict."""
    from . import synthetic
    from .grid import OccupancyGrid
    world = dist.get_world_size() if dist.is_initialized() else 1
    spec = synthetic.config(workload)
    N = perRank * world
    scene = synthetic.make_scene(seed=1, steps=steps, K=spec["K"], fov=spec["og"][4], unit=spec["og"][3])

    def run(cls):
        np.random.seed(seed)
        pf = cls(N, spec["og"], spec["sm"], device=dev)
        loc = pf.local if hasattr(pf, "local") else pf
        og = OccupancyGrid(*loc.geom.args, _geometry=loc.geom)
        for fr in scene["warm"]:
            og.updateOccupancyGrid(fr)
        loc.load_grid(og.device_grid)
        log = []
        for count, fr in enumerate(scene["frames"][:steps], start=1):
            pf.updateParticles(fr, count)
            fired = pf.weightUnbalanced()
            w = (pf._w if hasattr(pf, "local") else pf.weights).cpu().numpy().copy()
            log.append((fired, pf.poses().copy(), w))
            if count == 6:                       # force a resample to exercise the cross-rank lattice moves
                pf.resample()
        return pf, log

    state = np.random.get_state()
    try:
        spf, a = run(ShardedParticleFilter)
        ref, b = run(ParticleFilter)
    finally:
        np.random.set_state(state)
    ok = all(fa == fb and np.array_equal(pa, pb) and np.array_equal(wa, wb) for (fa, pa, wa), (fb, pb, wb) in zip(a, b))
    ok = ok and np.array_equal(spf.lastResampleIdx, ref.lastResampleIdx)
    for i in range(spf.lo, spf.hi):              # lattices after the forced resample + 3 more steps
        ok = ok and bool(torch.equal(spf.local.lattice(i - spf.lo), ref.lattice(i)))
    t = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t.item()) == 1
