// OccupancyGrid.updateOccupancyGrid for a batch of particles (Utils/OccupancyGrid.py:127-152), sm_100a.
//
// The reference walks the 180 beams; for beam i it takes the pre-binned list of lidar-local cells whose bearing
// sector is (start + offset + i) % numSpokes and applies
//     total[empty] += 1                      empty:  range_i < maxRange and r < range_i - wall/2
//     visited[hit] += 2; total[hit] += 2     hit:    range_i - wall/2 < r < range_i + wall/2
// where the map index of a local cell is rint((pose + local - mapLim0)/unit) and numpy's fancy `+=` applies
// once per distinct map cell per statement.
//
// Here the loop is inverted (owner computes): sectors of distinct beams are disjoint, so a local cell belongs to at
// most one beam.  A per-particle preparation pass evaluates the float64 index maps of the L local columns / rows;
// when both are pure shifts (always true for poses emitted by the matcher, which sit on the lattice) every local
// cell owns exactly one map cell and the fast path runs; otherwise (pose exactly half a cell off the lattice, where
// rint's half-to-even collapses neighbours) the general map-cell-owned path reproduces the per-statement union
// semantics.  Three launches per call: prep -> group (particles by sector shift) -> apply (+ general path).
#include <algorithm>
#include <cstdlib>

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
  // particles grouped by sector shift (update_group_kernel): all particles of a group see the same empty / hit flags
  int* hdr;           // [0] number of work items, [1] first sector of the union of the beam fans, [2] its length
  int* hist;          // [numSpokes] particles per sector shift, then the write cursor of each shift's group
  int4* items;        // work items (sectorShift, first slot in gbase / plist, particles (<= UPD_ITEM), all inside the map)
  long long* gbase;   // per grouped slot: cell index of the patch origin inside the lattice batch
  int* plist;         // per grouped slot: particle
  unsigned char* ginside;   // per grouped slot: the whole patch lies inside the lattice
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
  const double a0 = P.axis[0];
  const int sx = (int)rint(ddiv(dsub(dadd(x, a0), P.mapX0), P.unit));
  const int sy = (int)rint(ddiv(dsub(dadd(y, a0), P.mapY0), P.unit));
  bool pure = true;
  for (int l0 = lane; l0 < P.L; l0 += 32 * 8) {      // 8 independent axis loads in flight per lane
    double ax[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ax[j] = l0 + 32 * j < P.L ? __ldg(P.axis + l0 + 32 * j) : 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int l = l0 + 32 * j;
      if (l < P.L) {
        const int ix = (int)rint(ddiv(dsub(dadd(x, ax[j]), P.mapX0), P.unit));
        const int iy = (int)rint(ddiv(dsub(dadd(y, ax[j]), P.mapY0), P.unit));
        if (ix != sx + l || iy != sy + l) pure = false;
      }
    }
  }
  pure = __all_sync(0xffffffffu, pure);
  if (lane == 0) {
    // spokesOffsetIdxByTheta = int(rint(theta / (2*pi) * numSpokes))  (:131)
    const int off = (int)rint(dmul(ddiv(th, 6.283185307179586), (double)P.numSpokes));
    int shift = (P.start + off) % P.numSpokes;          // beam i looks at sector (shift + i) % numSpokes (:134)
    if (shift < 0) shift += P.numSpokes;
    P.prep[warp] = make_int4(sx, sy, shift, pure ? UPD_PURE : 0);
  }
}

__device__ __forceinline__ int beam_of(int sector, int shift, int numSpokes) {
  const int b = sector - shift;                  // inverse of spokeIdx = (start + off + i) % numSpokes (:134)
  return b < 0 ? b + numSpokes : b;              // sector, shift in [0, numSpokes)
}

// Particles grouped by sector shift.  All particles share the scan and differ only in the heading, quantised to whole
// spokes (:131): a population has a few dozen DISTINCT shifts, and every particle of a shift group sees the same empty /
// hit flag in every lidar-local cell.  One block: histogram over the shifts, compaction of the non-empty ones into work
// items of <= UPD_ITEM particles, particle slots in group order (patch origin inside the lattice batch, particle id).
constexpr int UPD_ITEM = 32;
constexpr int UPD_GROUP_T = 1024;
__global__ void __launch_bounds__(UPD_GROUP_T) update_group_kernel(UpdParams P) {
  __shared__ int s_scanA[UPD_GROUP_T / 32], s_scanB[UPD_GROUP_T / 32];
  __shared__ int s_fanMin, s_fanMax, s_slow;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nS = P.numSpokes;
  for (int s = tid; s < nS; s += UPD_GROUP_T) P.hist[s] = 0;
  if (tid == 0) { s_fanMin = 1 << 30; s_fanMax = -(1 << 30); s_slow = 0; }
  __syncthreads();
  // histogram + union of the beam fans as a signed circular distance to the first pure particle's shift
  int ref = -1;
  for (int p = 0; p < P.N && ref < 0; ++p)        // (uniform, normally one trip)
    if (P.prep[p].w & UPD_PURE) ref = P.prep[p].z;
  int dmin = 1 << 30, dmax = -(1 << 30);
  for (int p = tid; p < P.N; p += UPD_GROUP_T) {
    const int4 pr = P.prep[p];
    if (!(pr.w & UPD_PURE)) { P.slow[1 + atomicAdd(&s_slow, 1)] = p; continue; }     // general path (order irrelevant)
    atomicAdd(&P.hist[pr.z], 1);
    int d = pr.z - ref;
    d = ((d + nS / 2) % nS + nS) % nS - nS / 2;
    dmin = min(dmin, d); dmax = max(dmax, d);
  }
  dmin = __reduce_min_sync(0xffffffffu, dmin); dmax = __reduce_max_sync(0xffffffffu, dmax);
  if (lane == 0 && dmin <= dmax) { atomicMin(&s_fanMin, dmin); atomicMax(&s_fanMax, dmax); }
  __syncthreads();
  // exclusive scans over the spokes (a contiguous segment per thread): work items and particle slots before each spoke
  const int seg = (nS + UPD_GROUP_T - 1) / UPD_GROUP_T;
  const int s0 = min(tid * seg, nS), s1 = min(s0 + seg, nS);
  int nItems = 0, nPart = 0;
  for (int s = s0; s < s1; ++s) { const int c = P.hist[s]; nItems += (c + UPD_ITEM - 1) / UPD_ITEM; nPart += c; }
  int inclA = nItems, inclB = nPart;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, inclA, d), b = __shfl_up_sync(0xffffffffu, inclB, d);
    if (lane >= d) { inclA += a; inclB += b; }
  }
  if (lane == 31) { s_scanA[warp] = inclA; s_scanB[warp] = inclB; }
  __syncthreads();
  int baseA = 0, baseB = 0, totalA = 0;
  for (int w = 0; w < UPD_GROUP_T / 32; ++w) {
    if (w < warp) { baseA += s_scanA[w]; baseB += s_scanB[w]; }
    totalA += s_scanA[w];
  }
  int item = baseA + inclA - nItems, slot = baseB + inclB - nPart;
  for (int s = s0; s < s1; ++s) {
    const int c = P.hist[s];
    P.hist[s] = slot;                              // write cursor of this shift's group
    for (int q = 0; q < c; q += UPD_ITEM) P.items[item++] = make_int4(s, slot + q, min(UPD_ITEM, c - q), 1);
    slot += c;
  }
  if (tid == 0) {
    P.slow[0] = s_slow;
    P.hdr[0] = totalA;
    P.hdr[1] = s_fanMin <= s_fanMax ? ((ref + s_fanMin) % nS + nS) % nS : 0;
    P.hdr[2] = s_fanMin <= s_fanMax ? P.K + (s_fanMax - s_fanMin) : 0;
  }
  __syncthreads();
  const long long gstride = (long long)P.G * P.pitch;
  for (int p = tid; p < P.N; p += UPD_GROUP_T) {
    const int4 pr = P.prep[p];
    if (!(pr.w & UPD_PURE)) continue;
    const int q = atomicAdd(&P.hist[pr.z], 1);     // order inside a group is irrelevant: distinct particles, distinct lattices
    P.gbase[q] = (long long)lattice_of(P, p) * gstride + (long long)pr.y * P.pitch + pr.x;
    P.plist[q] = p;
    P.ginside[q] = (pr.x >= 0 && pr.y >= 0 && pr.x + P.L <= P.G && pr.y + P.L <= P.G) ? 1 : 0;
  }
}

// General path (non-pure index maps): thread owns a MAP cell of the patch's bounding box and tests the <= 3x3 local
// cells that can round onto it; union within a beam, sum across beams.
__device__ void update_general_one(const UpdParams& P, int p, int* mx, int* my);

// Fast path: a thread owns a lidar-local cell; a block walks over work items (a sector shift + <= UPD_ITEM particles
// that share it).  The cell's empty / hit flag is evaluated once per item; the particle loop is nothing but
// fire-and-forget float reductions at the L2 (RED.ADD.F32 / .F32x2), addresses = the particle's patch origin + the
// cell's offset.  The counts are small integers, so the float adds are exact, and every cell of a lattice is owned by
// one thread: the result does not depend on the order.  (Measured on c3: read-modify-write with 8 loads in flight per
// thread 104 -> reductions 90 us; the chunked round-2 kernel that re-evaluated the flags per 32-particle chunk: 144 us.)
// The blocks of row blockIdx.y == 0 first serve the particles of the general path (normally none).
constexpr int UPD_NB = 8;
__global__ void __launch_bounds__(256) update_apply_kernel(UpdParams P) {
  extern __shared__ double s_tab[];     // per BEAM: seen empty iff r < loE[b], hit iff lo[b] < r < hi[b]  (:138-143)
  if (blockIdx.y == 0) {
    const int count = P.slow[0];
    int* maps = reinterpret_cast<int*>(s_tab);       // mx[L], my[L]
    for (int q = 0; q < count; ++q) {
      update_general_one(P, P.slow[1 + q], maps, maps + P.L);
      __syncthreads();
    }
  }
  double* s_loE = s_tab;
  double* s_lo = s_tab + P.K;
  double* s_hi = s_tab + 2 * P.K;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  const int nS = P.numSpokes;
  int sec = 0;
  double r = 0.0;
  bool alive = cell < P.L * P.L;
  if (alive) {   // cells no particle can touch: outside the union of the beam fans, or beyond the reach of the scan
    sec = P.sector[cell];
    int rel = sec - P.hdr[1];
    if (rel < 0) rel += nS;
    alive = rel < P.hdr[2];
    if (alive) {
      r = P.radius[cell];
      alive = r < P.reach[0];
    }
  }
  if (!__syncthreads_or(alive)) return;
  for (int k = threadIdx.x; k < P.K; k += blockDim.x) {
    const double rm = P.ranges[k];
    const double lo = dsub(rm, P.wallHalf);
    s_lo[k] = lo;
    s_hi[k] = dadd(rm, P.wallHalf);
    s_loE[k] = rm < P.maxRange ? lo : -INFINITY;
  }
  __syncthreads();
  if (!alive) return;
  const int ly = cell / P.L, lx = cell - ly * P.L;
  const long long cellOff = (long long)ly * P.pitch + lx;
  float2* const grid = reinterpret_cast<float2*>(P.grid);
  const int nItems = P.hdr[0];
  for (int it = blockIdx.y; it < nItems; it += gridDim.y) {
    const int4 item = P.items[it];
    const int beam = beam_of(sec, item.x, nS);
    if (beam >= P.K) continue;
    const bool e = r < s_loE[beam], h = r > s_lo[beam] && r < s_hi[beam];
    if (!e && !h) continue;
    const long long* gb = P.gbase + item.y;
    const unsigned char* gi = P.ginside + item.y;
    const int m = item.z;
    for (int q0 = 0; q0 < m; q0 += UPD_NB) {
      float2* ptr[UPD_NB];
#pragma unroll
      for (int j = 0; j < UPD_NB; ++j) {
        ptr[j] = nullptr;
        if (q0 + j < m) {
          const long long b = gb[q0 + j];
          bool ok = true;
          if (!gi[q0 + j]) {      // patch not entirely inside the lattice: check this cell
            const int p = P.plist[item.y + q0 + j];
            const int4 pr = P.prep[p];
            const int jx = pr.x + lx, jy = pr.y + ly;
            ok = jx >= 0 && jy >= 0 && jx < P.G && jy < P.G;
            if (!ok) atomicOr(&P.status[p], SLAM_ST_SCAN_OUTSIDE_MAP);
          }
          if (ok) ptr[j] = grid + (b + cellOff);
        }
      }
#pragma unroll
      for (int j = 0; j < UPD_NB; ++j)
        if (ptr[j]) {
          if (h) atomicAdd(ptr[j], make_float2(2.f, 2.f));       // visited[hit] += 2; total[hit] += 2   (:151-152)
          else atomicAdd(&ptr[j]->y, 1.f);                       // total[empty] += 1                    (:149)
        }
    }
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
static size_t upd_al(size_t v) { return (v + 255) / 256 * 256; }
constexpr int UPD_MAX_SPOKES = 8192;
static size_t upd_items_cap(int32_t N) { return (size_t)N + (size_t)N / UPD_ITEM + 2; }   // <= one partial item per distinct shift

extern "C" size_t slam_update_workspace_bytes(int32_t N) {
  if (N <= 0) return 0;
  return upd_prep_bytes(N) + 256 + upd_al(((size_t)N + 1) * sizeof(int))              // prep records, reach, slow list
         + 256 + upd_al(UPD_MAX_SPOKES * sizeof(int)) + upd_al(upd_items_cap(N) * sizeof(int4))   // header, histogram, items
         + upd_al((size_t)N * sizeof(long long)) + upd_al((size_t)N * sizeof(int)) + upd_al((size_t)N) + 256;
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
  if (g->numSpokes > UPD_MAX_SPOKES) return fail(SLAM_E_UNSUPPORTED, "too many bearing sectors (numSpokes > 8192)");
  cudaStream_t st = (cudaStream_t)stream;
  // caller-provided scratch (no process-global state: several grids / streams / devices may call concurrently)
  if (!d_workspace || workspaceBytes < slam_update_workspace_bytes(N))
    return fail(SLAM_E_BADARG, "slam_update_grid: workspace too small (slam_update_workspace_bytes)");
  unsigned char* w = (unsigned char*)(((size_t)d_workspace + 255) / 256 * 256);
  auto take = [&](size_t bytes) { unsigned char* q = w; w += upd_al(bytes); return q; };
  int4* g_prep = reinterpret_cast<int4*>(take(upd_prep_bytes(N)));
  double* g_reach = reinterpret_cast<double*>(take(256));
  int* g_slow = reinterpret_cast<int*>(take(((size_t)N + 1) * sizeof(int)));
  UpdParams P;
  P.hdr = reinterpret_cast<int*>(take(256));
  P.hist = reinterpret_cast<int*>(take(UPD_MAX_SPOKES * sizeof(int)));
  P.items = reinterpret_cast<int4*>(take(upd_items_cap(N) * sizeof(int4)));
  P.gbase = reinterpret_cast<long long*>(take((size_t)N * sizeof(long long)));
  P.plist = reinterpret_cast<int*>(take((size_t)N * sizeof(int)));
  P.ginside = take((size_t)N);
  P.G = g->G; P.pitch = g->pitch; P.K = g->K; P.L = g->L; P.numSpokes = g->numSpokes; P.start = g->spokesStartIdx; P.N = N;
  P.unit = g->unit; P.mapX0 = g->mapX0; P.mapY0 = g->mapY0; P.maxRange = g->maxRange; P.wallHalf = g->wallHalf;
  P.sector = g->d_sector; P.radius = g->d_radius; P.axis = g->d_localAxis; P.ranges = d_ranges; P.pose = d_pose;
  P.grid = d_grid; P.prep = g_prep; P.slow = g_slow; P.status = d_status; P.slots = d_slots; P.reach = g_reach;
  update_prep_kernel<<<((N + 1) * 32 + 255) / 256, 256, 0, st>>>(P);      // + 1: the warp that computes the scan's reach
  SLAM_CUDA(cudaGetLastError());
  update_group_kernel<<<1, UPD_GROUP_T, 0, st>>>(P);
  SLAM_CUDA(cudaGetLastError());
  {
    // grid.x: 256-cell tiles of the local patch (of its (L+2)^2 bounding box for the general path); grid.y: the
    // blocks of a tile walk over the work items (<= N/32 + distinct shifts) with this stride -- small, because a tile
    // that no beam can touch is discovered (and left) once per block
    const int side = g->L + 2;
    const int ny = std::max(1, std::min(N / (4 * UPD_ITEM), 8));
    dim3 grid((side * side + 255) / 256, ny);
    const size_t smem = std::max(3 * (size_t)g->K * sizeof(double), 2 * (size_t)g->L * sizeof(int));
    update_apply_kernel<<<grid, 256, smem, st>>>(P);
    SLAM_CUDA(cudaGetLastError());
  }
  return 0;
}
