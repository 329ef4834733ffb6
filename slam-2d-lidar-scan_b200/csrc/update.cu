// OccupancyGrid.updateOccupancyGrid for a batch of particles (Utils/OccupancyGrid.py:127-152), sm_100a.
//
// The reference walks the 180 beams; for beam i it takes the pre-binned list of lidar-local cells whose bearing
// sector is (start + offset + i) % numSpokes and applies
//     total[empty] += 1                      empty:  range_i < maxRange and r < range_i - wall/2
//     visited[hit] += 2; total[hit] += 2     hit:    range_i - wall/2 < r < range_i + wall/2
// where the map index of a local cell is rint((pose + local - mapLim0)/unit) and numpy's fancy `+=` applies
// once per distinct map cell per statement.
//
// Here the loop is inverted (owner computes, no atomics): sectors of distinct beams are disjoint, so a local
// cell belongs to at most one beam.  A per-particle preparation pass evaluates the float64 index maps of the
// L local columns / rows; when both are pure shifts (always true for poses emitted by the matcher, which sit
// on the lattice) every local cell owns exactly one map cell and the fast kernel runs; otherwise (pose exactly
// half a cell off the lattice, where rint's half-to-even collapses neighbours) the general map-cell-owned
// kernel reproduces the per-statement union semantics.
#include "common.cuh"

namespace slam {

struct UpdParams {
  int G, pitch, K, L, numSpokes, start, N;
  double unit, mapX0, mapY0, maxRange, wallHalf;
  const short* sector;
  const double *radius, *axis, *ranges, *pose;
  float* grid;
  int4* prep;   // per particle: (shiftX, shiftY, spokeOffset, flags) flags: bit0 pure shift
  int* slow;    // slow[0] = number of particles needing the general path, slow[1..] = their indices
  int* status;
  const int* slots;   // physical lattice of particle p (null: p itself) -- copy-elided resampling keeps a slot table
  double* reach;      // [1] largest radius any beam of this scan can touch (empty or hit), written by the prep kernel
};

__device__ __forceinline__ size_t lattice_of(const UpdParams& P, int p) { return (size_t)(P.slots ? P.slots[p] : p); }

constexpr int UPD_PURE = 1;

// one warp per particle: float64 index maps of the local lattice (OccupancyGrid.py:144-145 via :102-106)
__global__ void update_prep_kernel(UpdParams P) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x >= blockDim.x - 32) {
    // last warp of the grid (spare by construction): the largest cell radius this scan can touch.  A beam empties
    // r < range - wall/2 (only if range < maxRange) and hits range - wall/2 < r < range + wall/2; the patch holds no
    // cell beyond sqrt(2) * maxRange, so a beam whose hit band starts past that (the log's 81.83 m sentinel) hits none.
    const double patchR = dmul(1.4143, P.maxRange);
    double reach = 0.0;
    for (int k = lane; k < P.K; k += 32) {
      const double rm = P.ranges[k], lo = dsub(rm, P.wallHalf), hi = dadd(rm, P.wallHalf);
      if (rm < P.maxRange) reach = fmax(reach, lo);
      if (lo < patchR) reach = fmax(reach, hi);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) reach = fmax(reach, __shfl_xor_sync(0xffffffffu, reach, d));
    if (lane == 0) P.reach[0] = reach;
    return;
  }
  if (warp >= P.N) return;
  const double x = P.pose[3 * warp], y = P.pose[3 * warp + 1], th = P.pose[3 * warp + 2];
  const int sx = (int)rint(ddiv(dsub(dadd(x, P.axis[0]), P.mapX0), P.unit));
  const int sy = (int)rint(ddiv(dsub(dadd(y, P.axis[0]), P.mapY0), P.unit));
  bool pure = true;
  for (int l = lane; l < P.L; l += 32) {
    const int ix = (int)rint(ddiv(dsub(dadd(x, P.axis[l]), P.mapX0), P.unit));
    const int iy = (int)rint(ddiv(dsub(dadd(y, P.axis[l]), P.mapY0), P.unit));
    if (ix != sx + l || iy != sy + l) pure = false;
  }
  pure = __all_sync(0xffffffffu, pure);
  if (lane == 0) {
    // spokesOffsetIdxByTheta = int(rint(theta / (2*pi) * numSpokes))  (:131)
    const int off = (int)rint(dmul(ddiv(th, 6.283185307179586), (double)P.numSpokes));
    int shift = (P.start + off) % P.numSpokes;          // beam i looks at sector (shift + i) % numSpokes (:134)
    if (shift < 0) shift += P.numSpokes;
    P.prep[warp] = make_int4(sx, sy, shift, pure ? UPD_PURE : 0);
    if (!pure) P.slow[1 + atomicAdd(&P.slow[0], 1)] = warp;
  }
}

__device__ __forceinline__ int beam_of(int sector, int shift, int numSpokes) {
  const int b = sector - shift;                  // inverse of spokeIdx = (start + off + i) % numSpokes (:134)
  return b < 0 ? b + numSpokes : b;              // sector, shift in [0, numSpokes)
}

// Fast path: thread owns a lidar-local cell and loops over a chunk of particles.  All particles share the scan; only
// the heading (quantised to whole spokes, :131) differs, so a chunk of 32 particles has a handful of DISTINCT sector
// shifts: the cell's empty / hit flags are evaluated once per distinct shift (2 bits each in a 64-bit mask) and the
// particle loop only does the read-modify-writes.  Blocks whose cells lie outside the union of the chunk's beam fans
// or beyond the reach of the scan leave before the interval tables are even built.
constexpr int UPD_CHUNK = 32;
__global__ void __launch_bounds__(256) update_fast_kernel(UpdParams P) {
  // per-sector interval tables of this scan: a cell of beam b is seen-empty iff r < loE[b] and hit iff lo[b] < r < hi[b]
  // (:138-143); sectors outside the fan get intervals that are never satisfied, so no beam < K test is needed
  extern __shared__ double s_tab[];
  double* s_loE = s_tab;
  double* s_lo = s_tab + P.numSpokes;
  double* s_hi = s_tab + 2 * P.numSpokes;
  __shared__ int4 s_prep[UPD_CHUNK];
  __shared__ size_t s_lat[UPD_CHUNK];
  __shared__ int s_fan[2];      // sectors any particle of the chunk can look at: [start, start + length) mod numSpokes
  __shared__ int s_dshift[UPD_CHUNK], s_didx[UPD_CHUNK], s_nd;     // distinct shifts, particle -> index into them
  const int p0 = blockIdx.y * UPD_CHUNK;
  const int np = min(UPD_CHUNK, P.N - p0);
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x, nS = P.numSpokes;
    int4 pr = make_int4(0, 0, 0, 0);
    if (lane < np) {
      pr = P.prep[p0 + lane];
      s_prep[lane] = pr;
      s_lat[lane] = lattice_of(P, p0 + lane);
    }
    // union of the particles' beam fans (headings of a chunk differ by a few spokes)
    const int ref = __shfl_sync(0xffffffffu, pr.z, 0);
    int d = lane < np ? pr.z - ref : 0;
    d = ((d + nS / 2) % nS + nS) % nS - nS / 2;        // signed circular distance to the first particle's shift
    const int dmin = __reduce_min_sync(0xffffffffu, d), dmax = __reduce_max_sync(0xffffffffu, d);
    if (lane == 0) { s_fan[0] = ((ref + dmin) % nS + nS) % nS; s_fan[1] = P.K + (dmax - dmin); }
    // distinct shifts: the lowest lane of every group of equal shifts is its leader
    const unsigned grp = __match_any_sync(0xffffffffu, lane < np ? pr.z : -1 - lane);
    const bool leader = lane < np && (__ffs(grp) - 1) == lane;
    const unsigned leaders = __ballot_sync(0xffffffffu, leader);
    if (leader) s_dshift[__popc(leaders & ((1u << lane) - 1u))] = pr.z;
    if (lane < np) s_didx[lane] = __popc(leaders & ((1u << (__ffs(grp) - 1)) - 1u));
    if (lane == 0) s_nd = __popc(leaders);
  }
  __syncthreads();
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  int sec = 0;
  double r = 0.0;
  bool alive = cell < P.L * P.L;
  if (alive) {   // cells no particle of the chunk can touch: outside the union of the fans, or beyond the reach of the scan
    sec = P.sector[cell];
    int rel = sec - s_fan[0];
    if (rel < 0) rel += P.numSpokes;
    alive = rel < s_fan[1];
    if (alive) {
      r = P.radius[cell];
      alive = r < P.reach[0];
    }
  }
  if (!__syncthreads_or(alive)) return;
  for (int k = threadIdx.x; k < P.numSpokes; k += blockDim.x) {
    double loE = -INFINITY, lo = INFINITY, hi = -INFINITY;
    if (k < P.K) {
      const double rm = P.ranges[k];
      lo = dsub(rm, P.wallHalf);
      hi = dadd(rm, P.wallHalf);
      if (rm < P.maxRange) loE = lo;
    }
    s_loE[k] = loE; s_lo[k] = lo; s_hi[k] = hi;
  }
  __syncthreads();
  if (!alive) return;
  const int ly = cell / P.L, lx = cell - ly * P.L;
  unsigned long long fm = 0ull;       // 2 bits per distinct shift: bit0 seen empty, bit1 hit
  const int nd = s_nd;
  for (int d = 0; d < nd; ++d) {
    const int beam = beam_of(sec, s_dshift[d], P.numSpokes);
    const unsigned long long f = (r < s_loE[beam] ? 1ull : 0ull) | ((r > s_lo[beam] && r < s_hi[beam]) ? 2ull : 0ull);
    fm |= f << (2 * d);
  }
  if (fm == 0ull) return;
  const size_t gstride = (size_t)P.G * P.pitch;
  int bad = 0;
  // particles of the chunk in groups of 8: all reads of a group are issued before its writes (distinct particles
  // own distinct lattices, so the read-modify-writes are independent)
  for (int q0 = 0; q0 < np; q0 += 8) {
    float2* ptr[8];
    float2 val[8];
    unsigned char flag[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      flag[j] = 0;
      ptr[j] = nullptr;
      const int q = q0 + j;
      if (q < np) {
        const int4 pr = s_prep[q];
        const unsigned char f = (unsigned char)((fm >> (2 * s_didx[q])) & 3ull);
        if (f && (pr.w & UPD_PURE)) {
          const int jx = pr.x + lx, jy = pr.y + ly;
          if (jx < 0 || jy < 0 || jx >= P.G || jy >= P.G) bad = 1;
          else {
            flag[j] = f;
            ptr[j] = (float2*)P.grid + s_lat[q] * gstride + (size_t)jy * P.pitch + jx;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (flag[j]) val[j] = *ptr[j];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (flag[j]) {
        float2 v = val[j];
        if (flag[j] & 1) v.y += 1.f;
        if (flag[j] & 2) { v.x += 2.f; v.y += 2.f; }
        *ptr[j] = v;
      }
  }
  if (bad) {
    for (int q = 0; q < np; ++q) {
      const int4 pr = s_prep[q];
      const int jx = pr.x + lx, jy = pr.y + ly;
      if ((pr.w & UPD_PURE) && ((fm >> (2 * s_didx[q])) & 3ull) && (jx < 0 || jy < 0 || jx >= P.G || jy >= P.G))
        atomicOr(&P.status[p0 + q], SLAM_ST_SCAN_OUTSIDE_MAP);
    }
  }
}

// General path (non-pure index maps): loops over the flagged particles; thread owns a MAP cell of the patch's
// bounding box and tests the <= 3x3 local cells that can round onto it; union within a beam, sum across beams.
__device__ void update_general_one(const UpdParams& P, int p, int* mx, int* my);

__global__ void __launch_bounds__(256) update_general_kernel(UpdParams P) {
  extern __shared__ int s_maps[];   // mx[L], my[L]
  const int count = P.slow[0];
  for (int q = 0; q < count; ++q) {
    update_general_one(P, P.slow[1 + q], s_maps, s_maps + P.L);
    __syncthreads();
  }
}

__device__ void update_general_one(const UpdParams& P, int p, int* mx, int* my) {
  const int4 pr = P.prep[p];
  const double x = P.pose[3 * p], y = P.pose[3 * p + 1];
  for (int l = threadIdx.x; l < P.L; l += blockDim.x) {
    mx[l] = (int)rint(ddiv(dsub(dadd(x, P.axis[l]), P.mapX0), P.unit));
    my[l] = (int)rint(ddiv(dsub(dadd(y, P.axis[l]), P.mapY0), P.unit));
  }
  __syncthreads();
  const int side = P.L + 2;   // bounding box in shifted map coordinates: [shift-1, shift+L]
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= side * side) return;
  const int cy = cell / side, cx = cell - cy * side;
  const int jx = pr.x - 1 + cx, jy = pr.y - 1 + cy;
  int nb = 0;
  int beams[9];
  unsigned char flag[9];   // bit0 empty, bit1 hit
  for (int dy = -1; dy <= 1; ++dy) {
    const int ly = cy - 1 + dy;
    if (ly < 0 || ly >= P.L || my[ly] != jy) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int lx = cx - 1 + dx;
      if (lx < 0 || lx >= P.L || mx[lx] != jx) continue;
      const int lc = ly * P.L + lx;
      const int beam = beam_of(P.sector[lc], pr.z, P.numSpokes);
      if (beam >= P.K) continue;
      const double rm = P.ranges[beam], r = P.radius[lc];
      const double lo = dsub(rm, P.wallHalf), hi = dadd(rm, P.wallHalf);
      const unsigned char f = ((rm < P.maxRange && r < lo) ? 1 : 0) | ((r > lo && r < hi) ? 2 : 0);
      if (!f) continue;
      int k = 0;
      for (; k < nb; ++k)
        if (beams[k] == beam) break;
      if (k == nb) { beams[nb] = beam; flag[nb] = 0; ++nb; }
      flag[k] |= f;
    }
  }
  if (nb == 0) return;
  float dv = 0.f, dt = 0.f;
  for (int k = 0; k < nb; ++k) {
    if (flag[k] & 1) dt += 1.f;
    if (flag[k] & 2) { dv += 2.f; dt += 2.f; }
  }
  if (jx < 0 || jy < 0 || jx >= P.G || jy >= P.G) { atomicOr(&P.status[p], SLAM_ST_SCAN_OUTSIDE_MAP); return; }
  float2* c = (float2*)P.grid + lattice_of(P, p) * P.G * P.pitch + (size_t)jy * P.pitch + jx;
  float2 v = *c;
  v.x += dv; v.y += dt;
  *c = v;
}

__global__ void grid_init_kernel(float4* g, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const float4 v = make_float4(1.f, 2.f, 1.f, 2.f);
  for (; i < n4; i += stride) g[i] = v;
}

}  // namespace slam

using namespace slam;

extern "C" int slam_grid_init(const slam_geometry* g, float* d_grid, int32_t N, void* stream) {
  if (!g || !d_grid || N < 0) return fail(SLAM_E_BADARG, "slam_grid_init: bad argument");
  const size_t n4 = (size_t)N * g->G * g->pitch / 2;
  if (n4 == 0) return 0;
  grid_init_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>((float4*)d_grid, n4);
  SLAM_CUDA(cudaGetLastError());
  return 0;
}

static size_t upd_prep_bytes(int32_t N) { return ((size_t)N * sizeof(int4) + 255) / 256 * 256; }

extern "C" size_t slam_update_workspace_bytes(int32_t N) {
  return N <= 0 ? 0 : upd_prep_bytes(N) + 256 + ((size_t)N + 1) * sizeof(int) + 256;
}

extern "C" int slam_update_grid(const slam_geometry* g, float* d_grid, int32_t N, const double* d_ranges,
                                const double* d_pose, int32_t* d_status, void* d_workspace, size_t workspaceBytes,
                                void* stream) {
  return slam_update_grid_slots(g, d_grid, nullptr, N, d_ranges, d_pose, d_status, d_workspace, workspaceBytes, stream);
}

extern "C" int slam_update_grid_slots(const slam_geometry* g, float* d_grid, const int32_t* d_slots, int32_t N,
                                      const double* d_ranges, const double* d_pose, int32_t* d_status, void* d_workspace,
                                      size_t workspaceBytes, void* stream) {
  if (!g || !d_grid || !d_ranges || !d_pose || !d_status) return fail(SLAM_E_BADARG, "slam_update_grid: null argument");
  if (N <= 0) return 0;
  if (g->K > SLAM_MAX_BEAMS) return fail(SLAM_E_UNSUPPORTED, "too many beams");
  cudaStream_t st = (cudaStream_t)stream;
  // caller-provided scratch (no process-global state: several grids / streams / devices may call concurrently):
  // [N] int4 preparation records + 1 + N ints of slow-path work list
  if (!d_workspace || workspaceBytes < slam_update_workspace_bytes(N))
    return fail(SLAM_E_BADARG, "slam_update_grid: workspace too small (slam_update_workspace_bytes)");
  const size_t prepBytes = upd_prep_bytes(N);
  unsigned char* scratch = (unsigned char*)(((size_t)d_workspace + 255) / 256 * 256);
  int4* g_prep = reinterpret_cast<int4*>(scratch);
  double* g_reach = reinterpret_cast<double*>(scratch + prepBytes);
  int* g_slow = reinterpret_cast<int*>(scratch + prepBytes + 256);
  SLAM_CUDA(cudaMemsetAsync(g_slow, 0, sizeof(int), st));
  UpdParams P;
  P.G = g->G; P.pitch = g->pitch; P.K = g->K; P.L = g->L; P.numSpokes = g->numSpokes; P.start = g->spokesStartIdx; P.N = N;
  P.unit = g->unit; P.mapX0 = g->mapX0; P.mapY0 = g->mapY0; P.maxRange = g->maxRange; P.wallHalf = g->wallHalf;
  P.sector = g->d_sector; P.radius = g->d_radius; P.axis = g->d_localAxis; P.ranges = d_ranges; P.pose = d_pose;
  P.grid = d_grid; P.prep = g_prep; P.slow = g_slow; P.status = d_status; P.slots = d_slots; P.reach = g_reach;
  update_prep_kernel<<<((N + 1) * 32 + 255) / 256, 256, 0, st>>>(P);      // + 1: the warp that computes the scan's reach
  SLAM_CUDA(cudaGetLastError());
  {
    dim3 grid((g->L * g->L + 255) / 256, (N + UPD_CHUNK - 1) / UPD_CHUNK);
    update_fast_kernel<<<grid, 256, 3 * g->numSpokes * sizeof(double), st>>>(P);
    SLAM_CUDA(cudaGetLastError());
  }
  {
    const int side = g->L + 2;
    dim3 grid((side * side + 255) / 256, 1);
    update_general_kernel<<<grid, 256, 2 * g->L * sizeof(int), st>>>(P);
    SLAM_CUDA(cudaGetLastError());
  }
  return 0;
}
