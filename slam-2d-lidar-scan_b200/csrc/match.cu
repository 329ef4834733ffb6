// Fused multi-resolution correlative scan matcher for a batch of particles (sm_100a).
//
// One persistent CTA per SM walks over particles; for each particle it runs the coarse and the fine stage of
// ScanMatcher.matchScan (Utils/ScanMatcher_OGBased.py:47-79) with every intermediate on chip, except the sparse
// fine likelihood field and the occupancy bitmap of the particle's window, which live in a per-CTA L2-resident slot:
//
//   union    one streaming read of the union of the coarse and all possible fine windows -> packed occupancy bits
//            (visited/total > 0.5 <=> 2*visited > total), issued one particle ahead and staggered across CTAs  (:29-31)
//   scatter  per stage: set bits -> float64 index maps -> row-major + transposed shared-memory bitmaps       (:32-37)
//   blur     separable symmetric correlation in scipy's exact operation order, exploiting that the input takes two
//            values {log(missProb), 0}: first pass = table lookup on a column's 2r+1 bits, second pass only on
//            active 32-cell tiles; background = host-computed constant from the same operation order          (:41-42)
//   clamp    probMin = global min (= the background constant whenever one inactive cell exists); applied in the blur (:43-44)
//   points   beam end points, compaction of beams < maxRange                                                  (:81-89)
//   lists    per theta: rotate, truncate to indices, warp bitonic sort + unique (lexicographic (x, y))        (:116-121)
//   scores   per (theta, dy, 2 adjacent dx): gathers + numpy-pairwise sum + priors                            (:125-132)
//   select   first-max argmax, or exp / pairwise sum / CDF inversion of one host uniform; confidence          (:133-141)
//
// Numerics: IEEE float64, no FMA contraction, numpy's pairwise-summation order and scipy's pair-add/multiply/
// accumulate order are reproduced literally so that the score volume is bit-identical to the oracle's.
#include <cuda.h>   // CUtensorMap (types only; the encoder is resolved through the runtime, no libcuda link)

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace slam {

#ifndef SLAM_STREAM_WARPS
#define SLAM_STREAM_WARPS 2
#endif
constexpr int NSW = SLAM_STREAM_WARPS;   // stream warps: TMA producers / occupancy-bit packers of the union window.  The
                                 // shared-memory pipe serves warps round-robin, so a stream warp reads its ring at 1/16
                                 // of the pipe (measured): NSW sets the stream rate (2.5 MB per particle at c3)
#ifndef SLAM_CTA_THREADS
#define SLAM_CTA_THREADS 512
#endif
constexpr int NT_ALL = SLAM_CTA_THREADS;   // 512 threads x 128 registers (4 warps per scheduler)
constexpr int NT = NT_ALL - 32 * NSW;   // compute threads per CTA
constexpr int NW = NT / 32;      // compute warps per CTA
constexpr int NWC = NW;          // compute warps
constexpr int NTC = NWC * 32;    // compute threads
constexpr unsigned FULL = 0xffffffffu;
constexpr int RING_STAGES = NSW > 2 ? 2 * NSW : 4;   // TMA ring depth over all stream warps (stages of ringRows window rows each)
constexpr int RING_PER_WARP = RING_STAGES / NSW;
static_assert(RING_STAGES % NSW == 0 && RING_PER_WARP >= 1, "ring stages must split evenly over the stream warps");

// Warp 0 is the stream warp (lowest warp id: the schedulers favour older warps, and the stream must never starve);
// compute threads are numbered from 0 by ctid().
__device__ __forceinline__ int ctid() { return (int)threadIdx.x - 32 * NSW; }

// barrier over the compute warps (named barrier 1)
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, %0;" ::"n"(NTC) : "memory"); }
__device__ __forceinline__ int csync_or(int pred) {
  int r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, 1, %2, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "r"(pred), "n"(NTC)
      : "memory");
  return r;
}

struct StageDev {
  double unitLength, logMiss;
  int r;                     // blur radius
  double w[2 * SLAM_MAX_BLUR_RADIUS + 1];
  double T1[SLAM_MAX_BLUR_RADIUS], T2[SLAM_MAX_BLUR_RADIUS];  // (c)*w[jj], (c+c)*w[jj]
  double C0;                 // c*w[r]
  double B1, B2;             // all-background value after the first / second pass
  const double* lutV;        // first-pass value for every (2r+1)-bit column pattern (built once, same op order)
  int nHalf, nOff, nTheta, nPoses;
  const double *thetas, *cosT, *sinT;
  int nLeaves;               // pairwise-sum leaves of the flattened score volume
  const int2* leaves;        // (offset, length)
  int nOps, nLevels;         // combine tree of the leaves, ops grouped by height: node[nLeaves + o] = node[a] + node[b]
  const int2* ops;
  int levelStart[24];
  // ---- plan
  int Wmax, Wmap, words, WT, Ppitch, TB, Kpad, E;   // words includes >= 1 always-zero spare word per row
  int nGrpPad;               // score tasks per offset row (>= ceil(nOff / GRP); padded so that dense gathers are bank-conflict free)
  int bitsInSmem, PInSmem, scoresInSmem, needScores;
  int oBits, oBitsT, oDil, oVw, oRow, oCol, oTiles, oP, oLists, oCnt, oScores, oDx, oDy, oLeaf;
  int tileCap;                 // capacity of one warp's private segment of the active-tile list
  int oAux, auxPitch, auxOK;   // range-path tables (4 x auxPitch shorts) + row buffer [Wmax][UW] words, if they fit
  size_t gBits, gBitsT, gDil, gTiles, gP, gScores;  // byte offsets inside a CTA's global scratch slot
};

struct MatchParams {
  int G, pitch, K, N;
  double unit, mapX0, mapX1, mapY0, mapY1, fovHalf, maxRange, R;
  const double *gridX, *gridY;
  StageDev st[2];
  const float* grid;
  const int* slots;      // physical lattice of particle p (null: p itself); see slam_copy_lattices
  const double *ranges, *estPose, *rv, *tw, *uniforms;
  double *outPose, *outConf;
  int *outIdx, *status;
  unsigned char* scratch;
  size_t slotBytes;
  double* dbgProb[2];
  int* dbgDims[2];
  double* dbgVol[2];
  long long* dbgCycles;  // [gridDim][48] or null
  int fieldOnlyStage;    // -1: the whole matchScan; 0 / 1: only build (and dump) that stage's likelihood field around estPose
  int forceExactCdf;
  int forceGenericScatter;   // test hook: always take the atomicOr scatter path
  int noPrune;               // test hook: evaluate every fine hypothesis completely (no branch-and-bound)
  int fast;              // plan: 1 = shared-memory plan for both stages, 0 = global-slot plan
  // background window stream (union of the coarse and every possible fine window, read once)
  double RU;             // union half size = R + coarse search radius + margin
  int UWcells;           // max columns of a union row segment (even)
  int UW;                // bitmap words per union row = 2*ceil(UWcells/64) (even/odd bit planes per 64 cells)
  int URows;             // max union rows
  int ringRows, oRing;   // window rows per TMA ring stage, shared-memory offset of the RING_STAGES-stage ring
  int boxCells, nBoxes;  // TMA box = boxCells cells x ringRows rows; nBoxes boxes side by side cover UWcells
  size_t gU;             // offset of the two union bitmaps inside a CTA's global slot
};

// ------------------------------------------------------------------------------------------------ device
__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -1 - i : (i >= n ? 2 * n - 1 - i : i); }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned lds_u32(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 lds_u32x4(unsigned a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ double lds_f64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}


// ---- mbarrier / TMA primitives (PTX; SASS: SYNCS.*, UTMALDG)
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n"
      "D_%=:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
// one box of the lattice tensor (cells as 64-bit elements; x, row, particle) -> shared memory, completion on `bar`;
// out-of-range elements arrive as zero (visited = total = 0 -> not occupied); streamed through L2 (evict first)
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0),
      "r"(c1), "r"(c2), "l"(0x12F0000000000000ull)
      : "memory");
}

// Buffers of the FAST plan are carved out of the dynamic shared-memory array; deriving them from the array itself
// (instead of a run-time select between shared and global) lets the compiler emit LDS/STS/ATOMS with 32-bit
// addressing instead of generic loads.  The SLOW plan (windows too large for shared memory) uses the global slot.
template <class T>
__device__ __forceinline__ T* sbuf(int off) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  return reinterpret_cast<T*>(smem_dyn + off);
}
template <bool FAST, class T>
__device__ __forceinline__ T* buf(int off, unsigned char* gslot, size_t goff) {
  if (FAST) return sbuf<T>(off);
  return reinterpret_cast<T*>(gslot + goff);
}

constexpr int GRP = 2;   // adjacent x offsets handled by one thread (share the key / bitmap lookups)

// Dense field (coarse stage, shared memory): GRP consecutive cells of one row per point.  32-bit shared addressing.
struct FetchDense {
  static constexpr bool kWide = false;   // shared-memory gathers: 8 points (16 loads) in flight are enough
  static constexpr bool kKeys8 = true;   // the phase is bound by shared-memory wavefronts: 8 keys in two 16-byte loads
  unsigned listS;        // shared address of the sorted unique keys (x << 16 | y)
  unsigned baseS;        // shared address of the field
  unsigned off;          // (dx << 16) + dy, added to the key in one go
  int pitch;
  // keys of points k .. k+7 (k a multiple of 8: two 16-byte broadcast loads instead of eight 4-byte ones)
  __device__ __forceinline__ void keys8(int k, unsigned (&key)[8]) const {
    const uint4 a = lds_u32x4(listS + 4u * k), b = lds_u32x4(listS + 4u * k + 16u);
    key[0] = a.x; key[1] = a.y; key[2] = a.z; key[3] = a.w; key[4] = b.x; key[5] = b.y; key[6] = b.z; key[7] = b.w;
  }
  __device__ __forceinline__ void get_key(unsigned key, double (&v)[GRP]) const {
    const unsigned sxy = key + off;
    const unsigned a = baseS + 8u * ((sxy & 0xffffu) * pitch + (sxy >> 16));
#pragma unroll
    for (int g = 0; g < GRP; ++g) v[g] = lds_f64(a + 8u * g);
  }
  __device__ __forceinline__ void get8(int k, double (&v)[8][GRP]) const {
    unsigned key[8];
    keys8(k, key);
#pragma unroll
    for (int l = 0; l < 8; ++l) get_key(key[l], v[l]);
  }
  __device__ __forceinline__ void get(int k, double (&v)[GRP]) const { get_key(lds_u32(listS + 4u * k), v); }
};

// Sparse field: materialised (already clamped) only where the activity bitmap is set; everywhere else it equals the
// all-background constant B2 (produced on the host by the same operation order as the blur).  Branch-free
// (predicated loads) so that the eight gathers of one pairwise step are all in flight together.
template <bool FAST>
struct FetchGated {
  static constexpr bool kWide = true;    // L2-latency gathers: 16 points (32 predicated loads) in flight per thread
  static constexpr bool kKeys8 = false;
  const unsigned* list;
  const unsigned* dil;   // activity bitmap [rows][words], last word of every row always zero
  unsigned listS, dilS;  // shared addresses of the same (FAST plan)
  const double* P;       // field base (no offset applied), global
  double B2;
  unsigned off;          // (dx << 16) + dy
  int pitch, words;
  __device__ __forceinline__ void keys8(int k, unsigned (&key)[8]) const {
    uint4 a, b;
    if (FAST) {
      a = lds_u32x4(listS + 4u * k); b = lds_u32x4(listS + 4u * k + 16u);
    } else {
      a = *reinterpret_cast<const uint4*>(list + k); b = *reinterpret_cast<const uint4*>(list + k + 4);
    }
    key[0] = a.x; key[1] = a.y; key[2] = a.z; key[3] = a.w; key[4] = b.x; key[5] = b.y; key[6] = b.z; key[7] = b.w;
  }
  __device__ __forceinline__ void get(int k, double (&v)[GRP]) const {
    get_key(FAST ? lds_u32(listS + 4u * k) : list[k], v);
  }
  __device__ __forceinline__ void get_key(unsigned key, double (&v)[GRP]) const {
    unsigned w0, w1;
    const unsigned sxy = key + off;
    const unsigned xx = sxy >> 16, yy = sxy & 0xffffu;
    const unsigned wi = yy * words + (xx >> 5);
    if (FAST) {
      w0 = lds_u32(dilS + 4u * wi);
      w1 = lds_u32(dilS + 4u * wi + 4u);
    } else {
      w0 = dil[wi];
      w1 = dil[wi + 1];
    }
    const unsigned f = __funnelshift_r(w0, w1, xx & 31u);
    const double* q = P + (yy * pitch + xx);
#pragma unroll
    for (int g = 0; g < GRP; ++g) {
      double pv = B2;
      if ((f >> g) & 1u) pv = __ldca(q + g);
      v[g] = pv;
    }
  }
};

// Exact branch-and-bound for the argmax-only (fine) stage.  Field values are <= 0 and IEEE addition is monotone, so
// the pairwise tree over the CURRENT lane sums (plus the finished first half, if any) is an upper bound of the
// hypothesis' final score; once it is strictly below the best finished score of the block (the incumbent, shared
// memory), the hypothesis can neither win nor tie the first-maximum rule and its remaining gathers are skipped.
struct PruneCtx {
  unsigned bestS;      // shared address of the incumbent score (double, <= 0)
  double a1[GRP];      // finished first half of the top-level split
  int h1;              // a1 is valid
  int validMask;       // which of the GRP sums are real hypotheses
  int pts;             // profiling: points actually gathered (8 per loop trip)
};

// numpy pairwise_sum for GRP independent sums sharing the index stream: n <= 128 branch inline, recursion
// (split at n/2 rounded down to a multiple of 8, depth <= 3 for n <= 512) as nested loops around ONE leaf body.
template <bool PRUNE, class F>
__device__ __forceinline__ bool leaf_sum_g(const F& f, int off, int n, double (&out)[GRP], PruneCtx& pc) {
  if (n < 8) {
#pragma unroll
    for (int g = 0; g < GRP; ++g) out[g] = 0.0;
    for (int i = 0; i < n; ++i) {
      double v[GRP];
      f.get(off + i, v);
#pragma unroll
      for (int g = 0; g < GRP; ++g) out[g] = dadd(out[g], v[g]);
    }
    return false;
  }
  double r[8][GRP];
  if constexpr (F::kKeys8) {
    f.get8(off, r);
  } else {
#pragma unroll
    for (int l = 0; l < 8; ++l) f.get(off + l, r[l]);
  }
  const int m = n - (n & 7);
  int i = 8;
  auto bound_below_incumbent = [&]() {      // pairwise tree over the CURRENT lane sums: an upper bound of the final score
    const double inc = lds_f64(pc.bestS);
    bool all = true;
#pragma unroll
    for (int g = 0; g < GRP; ++g) {
      double x = dadd(dadd(dadd(r[0][g], r[1][g]), dadd(r[2][g], r[3][g])), dadd(dadd(r[4][g], r[5][g]), dadd(r[6][g], r[7][g])));
      if (pc.h1) x = dadd(pc.a1[g], x);
      if (((pc.validMask >> g) & 1) && !(x < inc)) all = false;
    }
    return all;
  };
  if (F::kWide) {
    for (; i + 16 <= m; i += 16) {
      double v[16][GRP];
#pragma unroll
      for (int l = 0; l < 16; ++l) f.get(off + i + l, v[l]);      // sixteen points in flight; added in numpy's order below
#pragma unroll
      for (int l = 0; l < 8; ++l) {
#pragma unroll
        for (int g = 0; g < GRP; ++g) r[l][g] = dadd(r[l][g], v[l][g]);
      }
#pragma unroll
      for (int l = 0; l < 8; ++l) {
#pragma unroll
        for (int g = 0; g < GRP; ++g) r[l][g] = dadd(r[l][g], v[8 + l][g]);
      }
      if (PRUNE) {
        pc.pts += 16;
        if (bound_below_incumbent()) return true;
      }
    }
  }
  for (; i < m; i += 8) {
    double v[8][GRP];
    if constexpr (F::kKeys8) {
      f.get8(off + i, v);
    } else {
#pragma unroll
      for (int l = 0; l < 8; ++l) f.get(off + i + l, v[l]);     // eight independent gathers in flight
    }
    if (PRUNE) pc.pts += 8;
#pragma unroll
    for (int l = 0; l < 8; ++l) {
#pragma unroll
      for (int g = 0; g < GRP; ++g) r[l][g] = dadd(r[l][g], v[l][g]);
    }
    if (PRUNE && (F::kWide || (i & 8))) {                         // every 16 points (every 8 in the tail of the wide loop)
      if (bound_below_incumbent()) return true;
    }
  }
#pragma unroll
  for (int g = 0; g < GRP; ++g)
    out[g] = dadd(dadd(dadd(r[0][g], r[1][g]), dadd(r[2][g], r[3][g])), dadd(dadd(r[4][g], r[5][g]), dadd(r[6][g], r[7][g])));
  for (int i = m; i < n; ++i) {
    double v[GRP];
    f.get(off + i, v);
#pragma unroll
    for (int g = 0; g < GRP; ++g) out[g] = dadd(out[g], v[g]);
  }
  return false;
}

__device__ __forceinline__ int pw_split(int n) {   // numpy: n2 = n / 2; n2 -= n2 % 8
  int n2 = n / 2;
  return n2 - (n2 % 8);
}

// returns true if the hypothesis group was pruned (out is then undefined)
template <bool PRUNE, class F>
__device__ __forceinline__ bool pairwise_g(const F& f, int n, double (&out)[GRP], PruneCtx& pc) {
  const int c1 = n > 128 ? 2 : 1;
  pc.h1 = 0;
  for (int i1 = 0; i1 < c1; ++i1) {
    const int s1 = c1 == 2 ? pw_split(n) : n;
    const int o1 = i1 ? s1 : 0, n1 = c1 == 2 ? (i1 ? n - s1 : s1) : n;
    const int c2 = n1 > 128 ? 2 : 1;
    double acc2[GRP];
    if (PRUNE && i1) {
#pragma unroll
      for (int g = 0; g < GRP; ++g) pc.a1[g] = out[g];
      pc.h1 = 1;
    }
    for (int i2 = 0; i2 < c2; ++i2) {
      const int s2 = c2 == 2 ? pw_split(n1) : n1;
      const int o2 = o1 + (i2 ? s2 : 0), n2 = c2 == 2 ? (i2 ? n1 - s2 : s2) : n1;
      const int c3 = n2 > 128 ? 2 : 1;
      double acc3[GRP];
      for (int i3 = 0; i3 < c3; ++i3) {
        const int s3 = c3 == 2 ? pw_split(n2) : n2;
        const int o3 = o2 + (i3 ? s3 : 0), n3 = c3 == 2 ? (i3 ? n2 - s3 : s3) : n2;
        double leaf[GRP];
        if (leaf_sum_g<PRUNE>(f, o3, n3, leaf, pc)) return true;          // n3 <= 128 for n <= 512
#pragma unroll
        for (int g = 0; g < GRP; ++g) acc3[g] = i3 ? dadd(acc3[g], leaf[g]) : leaf[g];
      }
#pragma unroll
      for (int g = 0; g < GRP; ++g) acc2[g] = i2 ? dadd(acc2[g], acc3[g]) : acc3[g];
    }
#pragma unroll
    for (int g = 0; g < GRP; ++g) out[g] = i1 ? dadd(out[g], acc2[g]) : acc2[g];
  }
  return false;
}

// bitonic sort of 32*E keys held E per lane (element index = lane*E + e), ascending
template <int E>
__device__ __forceinline__ void warp_sort(unsigned (&key)[E], int lane) {
  constexpr int N = 32 * E;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j < E) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if ((e & j) == 0) {
            const int e2 = e | j;
            const int i = lane * E + e;
            const bool up = (i & k) == 0;
            unsigned a = key[e], b = key[e2];
            unsigned lo = min(a, b), hi = max(a, b);
            key[e] = up ? lo : hi;
            key[e2] = up ? hi : lo;
          }
        }
      } else {
        const int lj = j / E;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          unsigned other = __shfl_xor_sync(FULL, key[e], lj);
          const int i = lane * E + e;
          const bool up = (i & k) == 0;
          const bool lower = (lane & lj) == 0;
          key[e] = (lower == up) ? min(key[e], other) : max(key[e], other);
        }
      }
    }
  }
}

// a / b for a divisor whose correctly rounded reciprocal rb = RN(1 / b) is known: q = RN(a * rb), r = a - b * q (exact, one
// FMA), RN(q + r * rb) is the correctly rounded quotient (Markstein 1990; a, b normal and far from the exponent limits,
// which holds for window coordinates in metres divided by a cell size): bit-identical to __ddiv_rn, without the
// reciprocal refinement (MUFU + 5 DFMA) and the slow-path branch that the intrinsic repeats for every point.
__device__ __forceinline__ double ddiv_rcp(double a, double b, double rb) {
  const double q = __dmul_rn(a, rb);
  const double r = __fma_rn(-b, q, a);
  return __fma_rn(r, rb, q);
}

// One theta: rotate the K0 end points, truncate to field indices, sort, unique -> list, count.
template <int E>
__device__ __forceinline__ void build_list(const double* dxs, const double* dys, int K0, double ox, double oy,
                                           double c, double s, double bx, double by, double ul, int nHalf, int Wx,
                                           int Wy, unsigned* list, int* cnt, int lane, int& status) {
  const double rul = ddiv(1.0, ul);
  unsigned key[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    int k = lane * E + e;
    unsigned kk = 0xffffffffu;
    if (k < K0) {
      double ddx = dxs[k], ddy = dys[k];
      // ScanMatcher.rotate :169-170 -- evaluated left to right
      double qx = dsub(dadd(ox, dmul(c, ddx)), dmul(s, ddy));
      double qy = dadd(dadd(oy, dmul(s, ddx)), dmul(c, ddy));
      int xi = (int)ddiv_rcp(dsub(qx, bx), ul, rul);   // :174-175 astype(int) truncates toward zero
      int yi = (int)ddiv_rcp(dsub(qy, by), ul, rul);
      if (xi - nHalf < 0 || xi + nHalf >= Wx || yi - nHalf < 0 || yi + nHalf >= Wy) {
        status |= SLAM_ST_INDEX_OUT_OF_FIELD;
        xi = min(max(xi, nHalf), Wx - 1 - nHalf);
        yi = min(max(yi, nHalf), Wy - 1 - nHalf);
      }
      kk = ((unsigned)xi << 16) | (unsigned)yi;
    }
    key[e] = kk;
  }
  warp_sort<E>(key, lane);
  // unique: keep element i if it differs from element i-1
  unsigned prevLast = __shfl_up_sync(FULL, key[E - 1], 1);
  int mine = 0;
  bool keep[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    unsigned prev = (e == 0) ? prevLast : key[e - 1];
    bool first = (e == 0 && lane == 0);
    keep[e] = key[e] != 0xffffffffu && (first || key[e] != prev);
    mine += keep[e] ? 1 : 0;
  }
  int incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int v = __shfl_up_sync(FULL, incl, d);
    if (lane >= d) incl += v;
  }
  int pos = incl - mine;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if (keep[e]) {
      list[pos++] = key[e];
    }
  }
  int total = __shfl_sync(FULL, incl, 31);
  if (lane == 0) *cnt = total;
}

// numpy's argmax order: the first maximum in C order, and a NaN beats every number (np.argmax returns the first NaN)
__device__ __forceinline__ bool first_max_better(double v, int i, double b, int bi) {
  if (i < 0) return false;
  if (bi < 0) return true;
  const bool vn = v != v, bn = b != b;
  if (vn || bn) return vn && (!bn || i < bi);
  return v > b || (v == b && i < bi);
}

struct BlockScratch {
  double dval[NWC];
  int ival[NWC];
  double bcast[4];
  double incumbent;                // branch-and-bound: best finished score of the fine stage so far
  double wbest[NWC];               // per-warp first maximum of the score volume so far (merged batch by batch) ...
  int wbestIdx[NWC];               // ... and its flat index; -1 = none yet
  int nanFlag;                     // a NaN score was seen
  int taskCounter;                 // next 32-task chunk of the score phase (warps fetch chunks dynamically)
  int ibcast[8];
  unsigned maskA[64], maskB[64];   // per union-bitmap word: window columns whose index-map offset is D / D-1
};

__device__ __forceinline__ double block_min(double v, BlockScratch& bs) {
  int lane = ctid() & 31, warp = ctid() >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, d));
  csync();
  if (lane == 0) bs.dval[warp] = v;
  csync();
  double r = bs.dval[0];
  for (int i = 1; i < NWC; ++i) r = fmin(r, bs.dval[i]);
  return r;
}


// bitwise OR over the compute warps (status words)
__device__ __forceinline__ int block_or(int v, BlockScratch& bs) {
  v = __reduce_or_sync(FULL, v);
  csync();
  if ((ctid() & 31) == 0) bs.ival[ctid() >> 5] = v;
  csync();
  int r = 0;
  for (int i = 0; i < NWC; ++i) r |= bs.ival[i];
  return r;
}

// sub-phase cycle accounting (profiling hook): thread 0 of the compute warps accumulates the time since the last mark
struct SubCyc {
  long long* slot;
  long long last;
  __device__ __forceinline__ void start(long long* s) { slot = s; if (slot && ctid() == 0) last = clock64(); }
  __device__ __forceinline__ void mark(int i) {
    if (slot && ctid() == 0) { const long long t = clock64(); slot[i] += t - last; last = t; }
  }
};

constexpr int TPI = 2;   // generic-radius path: active tiles a warp blurs at a time
constexpr int TG = 4;    // templated-radius path: tiles per warp iteration (8 lanes x 4 cells each in the second pass)
constexpr int VW_PER_WARP = 192;   // doubles of first-pass values per warp: max(TG * (32 + 2*8), TPI * 64)

// ---- separable blur (phase D) templated on the radius so every tap loop unrolls and its loads pipeline.
// RT == 0: generic run-time radius.  Returns (through refs) the running minimum and the number of active cells.
template <int RT, bool FAST, bool DENSE>
__device__ __noinline__ void blur_stage(const StageDev& S, unsigned char* gslot, int Pp, int Wx, int Wy, int* counter,
                                        double& mnOut, int& activeOut, double& thrOut, long long* subSlots) {
  SubCyc sc;
  sc.start(subSlots);
  __shared__ double s_w[2 * SLAM_MAX_BLUR_RADIUS + 1];
  __shared__ int s_active[NWC];
  const unsigned* bits = buf<FAST, unsigned>(S.oBits, gslot, S.gBits);
  const unsigned* bitsT = buf<FAST, unsigned>(S.oBitsT, gslot, S.gBitsT);
  unsigned* dil = buf<FAST, unsigned>(S.oDil, gslot, S.gDil);
  double* Pf = DENSE ? sbuf<double>(S.oP) : reinterpret_cast<double*>(gslot + S.gP);
  double* VwAll = sbuf<double>(S.oVw);
  unsigned short* tiles = buf<FAST, unsigned short>(S.oTiles, gslot, S.gTiles);   // ids of the active tiles (order irrelevant)
  const int tid = ctid(), lane = tid & 31, warp = tid >> 5;
  const int r = RT ? RT : S.r;
  const int words = S.words, WT = S.WT;
  if (tid <= 2 * r) s_w[tid] = S.w[tid];
  (void)counter;
  csync();
  int myActive = 0;
  // Every warp appends the active tiles of its rows to a private segment of the list and blurs them itself
  // (rows are dealt out round-robin, so the segments balance): no shared-memory atomics -- contended
  // same-address atomics cost ~64 cycles each and used to dominate this phase.
  unsigned short* myTiles = tiles + (size_t)warp * S.tileCap;
  int myCount = 0;
  // D1. activity bitmap: cell (i, j) is active iff an occupied cell lies within +-r rows and +-r columns
  //     (reflected taps always fall inside that span, so plain dilation is exact).
  if (RT > 0 && words <= 32) {
    // vertical part first, one thread per (strip of RS rows, bitmap word): the rows of the strip (+-r) are loaded once
    // and the (2r+1)-row window OR is built by doubling in registers -- ~2 loads + 5 ORs per word instead of 2r+1 each
    constexpr int RS = 24, W = 2 * (RT ? RT : 1) + 1, NIN = RS + W - 1;
    constexpr int CMAX = W >= 16 ? 16 : (W >= 8 ? 8 : (W >= 4 ? 4 : (W >= 2 ? 2 : 1)));
    const int nStrips = (Wy + RS - 1) / RS;
    for (int u = tid; u < nStrips * words; u += NTC) {
      const int s = u / words, w = u - s * words;
      const int i0 = s * RS;
      unsigned a[NIN];
#pragma unroll
      for (int j = 0; j < NIN; ++j) {
        const int row = i0 - (RT ? RT : 1) + j;
        a[j] = (row >= 0 && row < Wy) ? bits[row * words + w] : 0u;
      }
#pragma unroll
      for (int c = 1; c < CMAX; c <<= 1) {
#pragma unroll
        for (int j = 0; j + c < NIN; ++j) a[j] |= a[j + c];      // ascending j: a[j + c] still holds the previous level
      }
#pragma unroll
      for (int j = 0; j < RS; ++j)
        if (i0 + j < Wy) dil[(i0 + j) * words + w] = a[j] | a[j + W - CMAX];
    }
    csync();
  }
  if (words <= 32) {
    const int seg = words <= 16 ? 16 : 32;               // lanes per row
    const int rowsPerWarp = 32 / seg;
    const int sub = lane / seg, w = lane - sub * seg;
    for (int i0 = warp * rowsPerWarp; i0 < Wy; i0 += NWC * rowsPerWarp) {
      const int i = i0 + sub;
      unsigned v = 0u;
      if (RT > 0) {
        if (i < Wy && w < words) v = dil[i * words + w];   // vertical window OR from the pre-pass (in place below)
      } else if (i < Wy && w < words) {
        if (i - r >= 0 && i + r < Wy) {                  // interior: no reflection, unit-stride rows
          const unsigned* q = bits + (i - r) * words + w;
          unsigned va = 0u, vb = 0u, vc = 0u, vd = 0u;     // four OR chains instead of one
#pragma unroll
          for (int d = 0; d <= 2 * r; d += 4) {
            va |= q[d * words];
            if (d + 1 <= 2 * r) vb |= q[(d + 1) * words];
            if (d + 2 <= 2 * r) vc |= q[(d + 2) * words];
            if (d + 3 <= 2 * r) vd |= q[(d + 3) * words];
          }
          v = (va | vb) | (vc | vd);
        } else {
#pragma unroll
          for (int d = -r; d <= r; ++d) v |= bits[reflect_idx(i + d, Wy) * words + w];
        }
      }
      unsigned lo = __shfl_up_sync(FULL, v, 1, seg), hi = __shfl_down_sync(FULL, v, 1, seg);
      if (w == 0) lo = 0u;
      if (w >= words - 1) hi = 0u;
      unsigned dl = v;
#pragma unroll
      for (int k = 1; k <= r; ++k) dl |= (v << k) | (v >> k) | (lo >> (32 - k)) | (hi << (32 - k));
      if (i < Wy && w < words) {
        const int valid = Wx - 32 * w;                   // columns of this word that exist
        if (valid < 32) dl &= valid <= 0 ? 0u : ((1u << valid) - 1u);
        dil[i * words + w] = dl;
        myActive += __popc(dl);
      }
      {   // warp-aggregated append of the active tiles of this step
        const bool act = (i < Wy && w < words) && dl != 0u;
        const unsigned am = __ballot_sync(FULL, act);
        if (act) myTiles[myCount + __popc(am & ((1u << lane) - 1u))] = (unsigned short)(i * words + w);
        myCount += __popc(am);
      }
    }
  } else {
    for (int t = tid; t < Wy * words; t += NTC) {
      const int i = t / words, w = t - i * words;
      unsigned lo = 0u, v = 0u, hi = 0u;
      for (int d = -r; d <= r; ++d) {
        const unsigned* row = bits + reflect_idx(i + d, Wy) * words;
        v |= row[w];
        if (w > 0) lo |= row[w - 1];
        if (w + 1 < words) hi |= row[w + 1];
      }
      unsigned dl = v;
      for (int k = 1; k <= r; ++k) dl |= (v << k) | (v >> k) | (lo >> (32 - k)) | (hi << (32 - k));
      const int valid = Wx - 32 * w;
      if (valid < 32) dl &= valid <= 0 ? 0u : ((1u << valid) - 1u);
      dil[t] = dl;
      myActive += __popc(dl);
    }
    for (int t0 = warp * 32; t0 < Wy * words; t0 += NTC) {     // same words, warp-aggregated append
      const int t = t0 + lane;
      const bool act = t < Wy * words && dil[t] != 0u;
      const unsigned am = __ballot_sync(FULL, act);
      if (act) myTiles[myCount + __popc(am & ((1u << lane) - 1u))] = (unsigned short)t;
      myCount += __popc(am);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) myActive += __shfl_xor_sync(FULL, myActive, d);
  if (lane == 0) s_active[warp] = myActive;
  csync();
  int nActive = 0;
  for (int w2 = 0; w2 < NWC; ++w2) nActive += s_active[w2];
  sc.mark(4);      // D1 dilation + tile list
  // probMin (:43): every blurred value is >= the all-background value B2 (each operation is monotone in its inputs
  // and the background has the lowest inputs), so as soon as one inactive cell exists probMin == B2 and the clamp
  // threshold is known before the blur; otherwise the caller takes the minimum and clamps afterwards.
  const bool anyInactive = nActive < Wx * Wy;
  const double thr = anyInactive ? dmul(0.5, S.B2) : 1.0;   // field values are <= 0: 1.0 never clamps
  const double* __restrict__ lut = S.lutV;
  // D2. active tiles (32 columns x 1 row) from the compacted list, TPI per warp at a time.
  //     First pass (axis 0) depends only on the 2r+1 occupancy bits of a column:
  //       out = x[c]*w[r]; out += (x[c+j] + x[c-j])*w[j+r], j = -r..-1, with x in {log(missProb), 0}.
  double mn = 0.0;
  double* Vw = VwAll + warp * VW_PER_WARP;
  if (RT > 0) {
    // TG tiles per warp iteration.  All 2*TG first-pass patterns / table look-ups of a lane are in flight together;
    // the second pass maps 8 lanes x 4 adjacent cells onto each tile, so a lane reads its 4 + 2r first-pass values
    // with 16-byte loads and keeps the taps in registers.
    constexpr int VS = 32 + 2 * (RT ? RT : 1), NV = 4 + 2 * (RT ? RT : 1), R = RT ? RT : 1;
    double wreg[R + 1];
#pragma unroll
    for (int jj = 0; jj <= R; ++jj) wreg[jj] = s_w[jj];
    const int u2 = lane >> 3, q4 = (lane & 7) * 4;
    const unsigned patMask = (2u << (2 * R)) - 1u;
    const unsigned wordsMagic = (unsigned)((0x100000000ull + (unsigned)words - 1) / (unsigned)words);   // ceil(2^32 / words)
    const int nList = myCount;
    __syncwarp();
    for (int base = 0; base < nList; base += TG) {
      int ti[TG], tw[TG];
      unsigned dls[TG];
      bool interior = true;      // no tile of the group touches a reflected border (identical in every lane)
#pragma unroll
      for (int u = 0; u < TG; ++u) {
        const bool real = base + u < nList;
        const int tt = (int)myTiles[real ? base + u : base];               // padding repeats the group's first tile ...
        dls[u] = real ? dil[tt] : 0u;                                      // ... with no active cell
        ti[u] = (int)__umulhi((unsigned)tt, wordsMagic);      // tt / words (tt < 2^16: exact)
        tw[u] = tt - ti[u] * words;
        interior = interior && ti[u] >= R && ti[u] + R < Wy && 32 * tw[u] >= R && 32 * tw[u] + 31 + R < Wx;
      }
      // virtual columns 32w-r .. 32w+31+r of every tile: lane handles vc0 = 32w-r+lane and (lane < 2r) vc1 = vc0+32
      unsigned sr[TG][2];
      if (__all_sync(FULL, interior)) {
        // the column's 2r+1 rows straight out of the transposed bitmap; columns and rows need no reflection
#pragma unroll
        for (int u = 0; u < TG; ++u) {
          const int lo = ti[u] - R;
          const unsigned* colBits = bitsT + (32 * tw[u] - R + lane) * WT + (lo >> 5);
          const unsigned* colBits1 = colBits + (lane < 2 * R ? 32 * WT : 0);     // lanes >= 2r: value unused
          sr[u][0] = __funnelshift_r(colBits[0], colBits[1], lo & 31) & patMask;
          sr[u][1] = __funnelshift_r(colBits1[0], colBits1[1], lo & 31) & patMask;
        }
      } else {
#pragma unroll 1
        for (int uh = 0; uh < 2 * TG; ++uh) {
          const int u = uh >> 1, h = uh & 1;
          const int i = u == 0 ? ti[0] : (u == 1 ? ti[1] : (u == 2 ? ti[2] : ti[3]));
          const int twu = u == 0 ? tw[0] : (u == 1 ? tw[1] : (u == 2 ? tw[2] : tw[3]));
          const int col = reflect_idx(32 * twu - R + lane + 32 * h, Wx);
          const unsigned* colBits = bitsT + col * WT;
          unsigned pat;
          if (i - R >= 0 && i + R < Wy) {
            const int lo = i - R;
            pat = __funnelshift_r(colBits[lo >> 5], colBits[(lo >> 5) + 1], lo & 31) & patMask;
          } else {                              // top / bottom border: reflected rows, bit by bit
            pat = 0u;
            for (int d = -R; d <= R; ++d) {
              const int rr = reflect_idx(i + d, Wy);
              pat |= ((colBits[rr >> 5] >> (rr & 31)) & 1u) << (d + R);
            }
          }
#pragma unroll
          for (int u2b = 0; u2b < TG; ++u2b) {
            if (u2b == u) { if (h == 0) sr[u2b][0] = pat; else sr[u2b][1] = pat; }
          }
        }
      }
      double vv[TG][2];
#pragma unroll
      for (int u = 0; u < TG; ++u) {
        vv[u][0] = __ldg(lut + sr[u][0]);       // first-pass value of this column pattern
        vv[u][1] = __ldg(lut + sr[u][1]);
      }
#pragma unroll
      for (int u = 0; u < TG; ++u) {
        Vw[u * VS + lane] = vv[u][0];
        if (lane < 2 * R) Vw[u * VS + lane + 32] = vv[u][1];
      }
      __syncwarp();
      // second pass (axis 1): lane (u2, q4) owns cells 32w + q4 .. q4+3 of tile u2
      const unsigned dsel = u2 == 0 ? dls[0] : (u2 == 1 ? dls[1] : (u2 == 2 ? dls[2] : dls[3]));
      const int tisel = u2 == 0 ? ti[0] : (u2 == 1 ? ti[1] : (u2 == 2 ? ti[2] : ti[3]));
      const int twsel = u2 == 0 ? tw[0] : (u2 == 1 ? tw[1] : (u2 == 2 ? tw[2] : tw[3]));
      const unsigned act4 = (dsel >> q4) & 0xfu;
      if (act4) {
        double c[NV];
        const double2* c2 = reinterpret_cast<const double2*>(Vw + u2 * VS + q4);
#pragma unroll
        for (int m = 0; m < NV / 2; ++m) {
          const double2 t2 = c2[m];
          c[2 * m] = t2.x; c[2 * m + 1] = t2.y;
        }
        double* out = Pf + (size_t)tisel * Pp + 32 * twsel + q4;
        double val[4];                                  // four independent dependency chains interleave
#pragma unroll
        for (int e = 0; e < 4; ++e) val[e] = dmul(c[e + R], wreg[R]);
#pragma unroll
        for (int jj = 0; jj < R; ++jj) {
#pragma unroll
          for (int e = 0; e < 4; ++e) val[e] = dadd(val[e], dmul(dadd(c[e + jj], c[e + 2 * R - jj]), wreg[jj]));
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if ((act4 >> e) & 1u) {
            out[e] = val[e] > thr ? 0.0 : val[e];       // clamp (:44)
          }
        }
      }
      __syncwarp();
    }
  } else {
    const int nTiles = Wy * words;
    (void)nTiles;
  const unsigned patMask = (2u << (2 * r)) - 1u;
  const int nList = myCount;
  __syncwarp();
  for (int base = 0; base < nList; base += TPI) {
    {
      int tt[TPI];
      unsigned dls[TPI];
#pragma unroll
      for (int u = 0; u < TPI; ++u) {
        tt[u] = base + u < nList ? (int)myTiles[base + u] : -1;
        dls[u] = tt[u] >= 0 ? dil[tt[u]] : 0u;
      }
      int ti[TPI], tw[TPI];
#pragma unroll
      for (int u = 0; u < TPI; ++u) {
        ti[u] = tt[u] < 0 ? 0 : tt[u] / words;
        tw[u] = tt[u] < 0 ? 0 : tt[u] - ti[u] * words;
      }
      // virtual columns 32w-r .. 32w+31+r: lane handles vc0 = 32w-r+lane and (lane < 2r) vc1 = vc0+32
#pragma unroll
      for (int u = 0; u < TPI; ++u) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = ti[u];
          const int col = reflect_idx(32 * tw[u] - r + lane + 32 * h, Wx);
          const unsigned* colBits = bitsT + col * WT;
          unsigned sr;
          if (i - r >= 0 && i + r < Wy) {       // the column's 2r+1 rows straight out of the transposed bitmap
            const int lo = i - r;
            sr = __funnelshift_r(colBits[lo >> 5], colBits[(lo >> 5) + 1], lo & 31) & patMask;
          } else {                              // top / bottom border: reflected rows, bit by bit
            sr = 0u;
            for (int d = -r; d <= r; ++d) {
              const int rr = reflect_idx(i + d, Wy);
              sr |= ((colBits[rr >> 5] >> (rr & 31)) & 1u) << (d + r);
            }
          }
          const double v = __ldg(lut + sr);     // first-pass value of this column pattern
          if (h == 0 || lane < 2 * r) Vw[u * 64 + lane + 32 * h] = v;
        }
      }
      __syncwarp();
      // second pass (axis 1) for the active cells of the tiles
#pragma unroll
      for (int u = 0; u < TPI; ++u) {
        if ((dls[u] >> lane) & 1u) {
          const double* c = Vw + u * 64 + lane + r;       // virtual column of cell 32w+lane
          double val = dmul(c[0], s_w[r]);
#pragma unroll
          for (int jj = 0; jj < r; ++jj) val = dadd(val, dmul(dadd(c[jj - r], c[r - jj]), s_w[jj]));
          mn = fmin(mn, val);
          Pf[(size_t)ti[u] * Pp + 32 * tw[u] + lane] = val > thr ? 0.0 : val;     // clamp (:44)
        }
      }
      __syncwarp();
    }
  }
  }
  if (RT > 0 && !anyInactive) {
    // no background cell at all (tiny windows): probMin is not known in advance -- take it from the finished field
    csync();
    for (int i = tid; i < Wy * Wx; i += NTC) {
      const int yy = i / Wx, xx = i - yy * Wx;
      mn = fmin(mn, Pf[(size_t)yy * Pp + xx]);
    }
  }
  sc.mark(5);      // D2 (own tiles; the wait for the other warps is accounted to the caller)
  if (subSlots && !DENSE && lane == 0) {      // profiling: active tiles / cells of the fine field
    atomicAdd((unsigned long long*)subSlots + 14, (unsigned long long)myCount);
    if (warp == 0) atomicAdd((unsigned long long*)subSlots + 15, (unsigned long long)nActive);
  }
  mnOut = mn;
  activeOut = nActive;
  thrOut = anyInactive ? thr : 0.0;     // 0.0 = "not clamped yet"
}


struct ScoreArgs {
  unsigned char* gslot;
  const double *rv, *tw;
  double* dvol;
  double B2, thr;
  size_t gP, gDil, gScores;
  int oLists, oCnt, oP, oDil, oScores, needScores;
  int Kpad, Pp, words, nHalf, nOff, nt, t0, nGrpPad;
  unsigned bestS;       // shared address of the branch-and-bound incumbent (fine stage)
  double* bestP;        // the same as a pointer
  long long* ptsSlot;   // profiling: [0] += points gathered, [1] += points of the evaluated hypotheses without pruning
  BlockScratch* bs;     // per-warp running maxima / NaN flag (results leave through shared memory: by-reference
                        // results would be promoted to registers that are reserved along the whole call chain)
};

// Score volume of a batch of thetas (phase G2).  Kept out of line so the hot gather loop gets its own register
// allocation instead of inheriting the pressure of the rest of the fused kernel.
template <bool FAST, bool DENSE, bool PRUNE>
__device__ __noinline__ void score_batch(const ScoreArgs& A) {
  const unsigned* lists = sbuf<unsigned>(A.oLists);
  const int* cnts = sbuf<int>(A.oCnt);
  const double* Pf = DENSE ? sbuf<double>(A.oP) : reinterpret_cast<const double*>(A.gslot + A.gP);
  const unsigned* dil = buf<FAST, unsigned>(A.oDil, A.gslot, A.gDil);
  double* scores = A.needScores ? buf<FAST, double>(A.oScores, A.gslot, A.gScores) : nullptr;
  const int tid = ctid();
  const int nOff = A.nOff, nOff2 = nOff * nOff;
  const int nGrp = A.nGrpPad;        // tasks per offset row; tasks with b0 >= nOff are padding (idle lanes)
  const int perTheta = nOff * nGrp;
  const int nq = A.nt * perTheta;
  double best = 0.0;
  int bestIdx = -1, sawNan = 0;
  PruneCtx pc;
  pc.bestS = A.bestS; pc.h1 = 0; pc.validMask = 0; pc.pts = 0;
  int ptsAll = 0;
#pragma unroll
  for (int g = 0; g < GRP; ++g) pc.a1[g] = 0.0;
  if (PRUNE && A.t0 == 0) {
    // Warm start of the incumbent: every warp sums one hypothesis next to the centre of the search (the coarse
    // winner: rotations mid-1 .. mid+1, offsets 0 and +-1 cell) with its lanes striding over the points -- not in
    // numpy's order, so the sum minus a margin far above its rounding error (<= 512 * 2^-53 * |sum|) is used: a valid
    // lower bound of that hypothesis' exact score, hence of the maximum.  Hopeless hypotheses are then dropped at
    // their first bound check instead of running until the first exact score exists.
    const int lane = tid & 31, warp = tid >> 5, mid = A.nt >> 1;
    for (int c = warp; c < 15; c += NWC) {
      const int tl = min(max(mid + (c % 3) - 1, 0), A.nt - 1), o = c / 3;
      const int step = A.nHalf > 0 ? 1 : 0;
      const int da = step * ((o == 3) - (o == 4)), db = step * ((o == 1) - (o == 2));
      const unsigned off = (unsigned)((db << 16) + da);
      const unsigned* list = lists + tl * A.Kpad;
      const int n = cnts[tl];
      double sum = 0.0;
      for (int k = lane; k < n; k += 32) {
        const unsigned sxy = list[k] + off, xx = sxy >> 16, yy = sxy & 0xffffu;
        const bool on = (dil[yy * A.words + (xx >> 5)] >> (xx & 31u)) & 1u;
        sum += on ? __ldca(Pf + (yy * A.Pp + xx)) : A.B2;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(FULL, sum, d);
      if (lane == 0) {
        const double lb = sum - 1e-9 * (1.0 + fabs(sum));          // scores are <= 0: the larger double has the smaller bit pattern
        atomicMin(reinterpret_cast<unsigned long long*>(A.bestP), (unsigned long long)__double_as_longlong(lb));
      }
    }
    csync();
  }
  // Warps fetch 32-task chunks from a shared counter: with the branch-and-bound a chunk next to the optimum costs
  // several times a pruned one, and a static split leaves warps waiting at the barrier that ends the phase.
  const int lane_ = tid & 31;
  for (;;) {
    int chunk = 0;
    if (lane_ == 0) chunk = atomicAdd(&A.bs->taskCounter, 1);
    chunk = __shfl_sync(FULL, chunk, 0);
    if (chunk * 32 >= nq) break;
    const int q = chunk * 32 + lane_;
    if (q >= nq) continue;
    int tl = q / perTheta;
    const int rem0 = q - tl * perTheta;
    if (PRUNE) {      // most promising rotations first (centre of the batch outwards): good incumbents early
      const int mid = A.nt >> 1;                        // bijection of [0, nt): mid, mid-1, mid+1, mid-2, ...
      tl = (tl & 1) ? mid - 1 - (tl >> 1) : mid + (tl >> 1);
    }
    const int a = rem0 / nGrp, b0 = (rem0 - a * nGrp) * GRP;
    if (b0 >= nOff) continue;
    double sc[GRP];
    bool pruned = false;
    if (DENSE) {
      FetchDense f;
      f.listS = smem_u32(lists + tl * A.Kpad);
      f.baseS = smem_u32(Pf);
      f.off = (unsigned)(((b0 - A.nHalf) << 16) + (a - A.nHalf));
      f.pitch = A.Pp;
      pairwise_g<false>(f, cnts[tl], sc, pc);                      // np.sum(axis=2) :130
    } else {
      FetchGated<FAST> f;
      f.list = lists + tl * A.Kpad;
      f.dil = dil;
      f.listS = smem_u32(f.list);
      f.dilS = FAST ? smem_u32(dil) : 0u;
      f.P = Pf; f.B2 = A.B2; f.pitch = A.Pp; f.words = A.words;
      f.off = (unsigned)(((b0 - A.nHalf) << 16) + (a - A.nHalf));
      pc.validMask = (b0 + 1 < nOff) ? 3 : 1;
      ptsAll += cnts[tl];
      pruned = pairwise_g<PRUNE>(f, cnts[tl], sc, pc);
    }
    if (pruned) continue;
#pragma unroll
    for (int g = 0; g < GRP; ++g) {
      const int b = b0 + g;
      if (b < nOff) {
        const int rem = a * nOff + b;
        double v = sc[g];
        if (A.rv) v = dadd(v, A.rv[rem]);                          // + rv + thetaWeight :131
        if (A.tw) v = dadd(v, A.tw[rem]);
        const int flat = (A.t0 + tl) * nOff2 + rem;
        if (scores) scores[flat] = v;
        if (A.dvol) A.dvol[flat] = v;
        if (v != v) sawNan = 1;
        if (first_max_better(v, flat, best, bestIdx)) { best = v; bestIdx = flat; }
        if (PRUNE && v > lds_f64(A.bestS)) {
          // scores are <= 0: the larger double has the smaller bit pattern
          atomicMin(reinterpret_cast<unsigned long long*>(A.bestP), (unsigned long long)__double_as_longlong(v));
        }
      }
    }
  }
  if (PRUNE && A.ptsSlot) {
    const int a = __reduce_add_sync(FULL, pc.pts), b = __reduce_add_sync(FULL, ptsAll);
    if ((tid & 31) == 0) { atomicAdd((unsigned long long*)A.ptsSlot, (unsigned long long)a); atomicAdd((unsigned long long*)A.ptsSlot + 1, (unsigned long long)b); }
  }
  // first maximum in C order over the warp, merged into the warp's running maximum of the earlier batches
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double ob = __shfl_xor_sync(FULL, best, d);
    const int oi = __shfl_xor_sync(FULL, bestIdx, d);
    if (first_max_better(ob, oi, best, bestIdx)) { best = ob; bestIdx = oi; }
  }
  sawNan = __any_sync(FULL, sawNan);
  if ((tid & 31) == 0) {
    const int warp = tid >> 5;
    const double ob = A.bs->wbest[warp];
    const int oi = A.bs->wbestIdx[warp];
    if (first_max_better(best, bestIdx, ob, oi)) { A.bs->wbest[warp] = best; A.bs->wbestIdx[warp] = bestIdx; }
    if (sawNan) A.bs->nanFlag = 1;
  }
}

struct ListArgs {
  const double *cosT, *sinT;
  double ox, oy, bx, by, ul;
  int oDx, oDy, oLists, oCnt;
  int K0, nHalf, Wx, Wy, Kpad, nt, t0;
};

template <int E>
__device__ __noinline__ void lists_batch(const ListArgs& A, int& statusIO) {
  const double* dxs = sbuf<double>(A.oDx);
  const double* dys = sbuf<double>(A.oDy);
  unsigned* lists = sbuf<unsigned>(A.oLists);
  int* cnts = sbuf<int>(A.oCnt);
  const int lane = ctid() & 31, warp = ctid() >> 5;
  int status = statusIO;
  for (int tl = warp; tl < A.nt; tl += NWC)
    build_list<E>(dxs, dys, A.K0, A.ox, A.oy, A.cosT[A.t0 + tl], A.sinT[A.t0 + tl], A.bx, A.by, A.ul, A.nHalf,
                  A.Wx, A.Wy, lists + tl * A.Kpad, cnts + tl, lane, status);
  statusIO = status;
}


// ---- union window ------------------------------------------------------------------------------------------
// The coarse window and every fine window the coarse stage can select lie inside est +- RU.  That union is read
// ONCE per particle (instead of once per stage): 16-byte streaming loads, two rows / 18 loads in flight per lane,
// thresholded visited/total > 0.5 <=> 2*visited > total (ScanMatcher_OGBased.py:29-31) and packed with warp ballots
// into an L2-resident bitmap (word 2t = cells 64t+2l, word 2t+1 = cells 64t+2l+1 of a row).  Each stage then
// scatters the set bits of its own window through its float64 index maps (:36-37).
__device__ __forceinline__ void union_window(const MatchParams& P, int p, int (&w)[4]) {
  const double x = P.estPose[3 * p], y = P.estPose[3 * p + 1];
  int x0 = (int)floor(ddiv(dsub(dsub(x, P.RU), P.mapX0), P.unit)) - 1;
  int y0 = (int)floor(ddiv(dsub(dsub(y, P.RU), P.mapY0), P.unit)) - 1;
  int x1 = (int)ceil(ddiv(dsub(dadd(x, P.RU), P.mapX0), P.unit)) + 2;
  int y1 = (int)ceil(ddiv(dsub(dadd(y, P.RU), P.mapY0), P.unit)) + 2;
  x0 = max(x0, 0) & ~1; y0 = max(y0, 0);
  x1 = min(x1, P.G); y1 = min(y1, P.G);
  int nc = max(x1 - x0, 0);
  nc = min((nc + 1) & ~1, P.UWcells);
  if (x0 + nc > P.pitch) nc = (P.pitch - x0) & ~1;
  w[0] = y0; w[1] = x0; w[2] = min(max(y1 - y0, 0), P.URows); w[3] = nc;
}

// Cold start: nothing overlaps the stream of a CTA's FIRST particle, so the compute warps pack most of its rows
// themselves with plain loads (rows >= first_rows_by_stream) while the stream warps take the head of the window.
__device__ __forceinline__ int first_rows_by_stream(int rows, int rowsPerChunk) {
  const int q = rowsPerChunk * NSW;
  return min(rows, max(q, (rows / 5) / q * q));      // measured: the stream warps pack ~0.2 of the rows in the same time
}

// mbarriers of the stream pipeline (static shared memory)
struct StreamShared {
  unsigned long long ringFull[RING_STAGES];   // TMA complete_tx: a ring stage has landed
  unsigned long long uFull[2], uEmpty[2];     // union bitmap b is complete / may be overwritten
};

// One window row of a ring stage -> NTW pairs of bitmap words (plain layout: word w = cells 32w .. 32w+31 of the row).
// All loads of the row are issued before the first ballot so that the single stream warp runs with full ILP.
__device__ __forceinline__ float2 lds_f2(unsigned a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
// 64-cell groups per TMA box for a row of n groups: the largest divisor of n that keeps a box <= 256 cells
__host__ __device__ constexpr int groups_per_box(int n) { return n % 4 == 0 ? 4 : (n % 3 == 0 ? 3 : (n % 2 == 0 ? 2 : 1)); }

// NR window rows at a time -> bitmap words (plain layout: word w of a row = cells 32w .. 32w+31).
// The shared-memory pipe is shared with the compute warps' gathers and serves warps round-robin, so a stream warp
// reads its ring at a fraction of the pipe: every load of the batch is issued before the first ballot.
template <int NTW, int NR>
__device__ __forceinline__ void pack_rows(unsigned rowS, int rowBytes, int boxRowBytes, unsigned* Urow, int UW, int nr, int lane) {
  constexpr int tPerBox = groups_per_box(NTW);
  float2 lo[NR][NTW], hi[NR][NTW];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
#pragma unroll
    for (int t = 0; t < NTW; ++t) {
      // rows beyond the batch re-read its first row (a stage may hold a single row): never past the ring
      const unsigned a = rowS + (unsigned)((r < nr ? r : 0) * rowBytes + (t / tPerBox) * boxRowBytes + (t % tPerBox) * 512 + lane * 8);
      lo[r][t] = lds_f2(a);            // cell 64t + lane       (visited, total)
      hi[r][t] = lds_f2(a + 256u);     // cell 64t + 32 + lane
    }
  }
  // every ballot is warp-uniform: lane 0 writes the row with 16-byte stores (UW is a multiple of 4 words; words past
  // the window are zero)
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    if (r < nr) {
      unsigned w[(2 * NTW + 3) / 4 * 4];
#pragma unroll
      for (int t = 0; t < (2 * NTW + 3) / 4 * 2; ++t) {
        if (t < NTW) {
          w[2 * t] = __ballot_sync(FULL, 2.f * lo[r][t].x > lo[r][t].y);       // visited/total > 0.5  (:29-31)
          w[2 * t + 1] = __ballot_sync(FULL, 2.f * hi[r][t].x > hi[r][t].y);
        } else {
          w[2 * t] = 0u; w[2 * t + 1] = 0u;
        }
      }
      if (lane == 0) {
        uint4* dst = reinterpret_cast<uint4*>(Urow + (size_t)r * UW);
#pragma unroll
        for (int g = 0; g < (2 * NTW + 3) / 4; ++g) dst[g] = make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
      }
    }
  }
}

// ---- stream warp.  For every particle of this CTA it stages the union window through a shared-memory ring with
// TMA tensor loads (cp.async.bulk.tensor, one elected lane; RING_STAGES stages of ringRows rows in flight, no
// registers tied up) and packs visited/total > 0.5 into the particle's union bitmap while the compute warps are
// still busy with the previous particle.  Out-of-lattice cells arrive as zeros (not occupied).
__device__ __noinline__ void stream_role(const MatchParams& P, const CUtensorMap* tmap, StreamShared& sh, int (*s_uwin)[4]) {
  const int lane = threadIdx.x & 31, sw = threadIdx.x >> 5;       // stream warp sw takes the row chunks c = sw (mod NSW)
  unsigned char* gslot = P.scratch + (size_t)blockIdx.x * P.slotBytes;
  unsigned* Ubuf0 = reinterpret_cast<unsigned*>(gslot + P.gU);
  const int BY = P.ringRows, BX = P.boxCells, nB = P.nBoxes, UW = P.UW;
  const unsigned boxBytes = (unsigned)(BY * BX * 8), stageBytes = boxBytes * nB;
  const unsigned ring = ((smem_u32(sbuf<unsigned char>(0)) + (unsigned)P.oRing + 127u) & ~127u) + sw * RING_PER_WARP * stageBytes;
  const int tPerBox = BX / 64, nTW = P.UWcells / 64;
  const int nMine = (P.N - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  long long* cyc = (P.dbgCycles && threadIdx.x == 0) ? P.dbgCycles + (size_t)blockIdx.x * 48 : nullptr;   // [7] TMA wait, [14] pack, [15] bitmap wait
  // producer cursor (warp-uniform): particle kP, own chunk cP (counts this warp's chunks), window wP
  int kP = 0, cP = 0, nChP = -1, latP = 0, wP[4] = {0, 0, 0, 0};
  unsigned issued = 0, consumed = 0;
  auto my_chunks = [&](int rows) { const int n = (rows + BY - 1) / BY; return n > sw ? (n - sw + NSW - 1) / NSW : 0; };
  auto top_up = [&]() {
    while (issued - consumed < (unsigned)RING_PER_WARP && kP < nMine) {
      const int pP = (int)blockIdx.x + kP * (int)gridDim.x;
      if (nChP < 0) {
        union_window(P, pP, wP);
        latP = P.slots ? P.slots[pP] : pP;
        nChP = my_chunks(kP == 0 ? first_rows_by_stream(wP[2], BY) : wP[2]);
        cP = 0;
      }
      if (cP >= nChP) { ++kP; nChP = -1; continue; }
      const unsigned st = issued % RING_PER_WARP;
      if (lane == 0) {
        const unsigned bar = smem_u32(&sh.ringFull[sw * RING_PER_WARP + st]);
        mbar_expect_tx(bar, stageBytes);
        for (int j = 0; j < nB; ++j)
          tma_load_3d(ring + st * stageBytes + j * boxBytes, tmap, bar, wP[1] + j * BX, wP[0] + (sw + cP * NSW) * BY, latP);
      }
      ++issued; ++cP;
    }
  };
  for (int k = 0; k < nMine; ++k) {
    const int p = (int)blockIdx.x + k * (int)gridDim.x, b = k & 1;
    int w[4];
    union_window(P, p, w);
    const int rowsMine = k == 0 ? first_rows_by_stream(w[2], BY) : w[2];     // rows this pipeline covers
    const int nCh = my_chunks(rowsMine);
    unsigned* U = Ubuf0 + (size_t)b * P.URows * UW;
    top_up();                                                     // loads of this (and the next) particle in flight
    if (cyc) cyc[15] -= clock64();
    mbar_wait(smem_u32(&sh.uEmpty[b]), (unsigned)(((k >> 1) & 1) ^ 1));   // compute is done with bitmap b
    if (cyc) cyc[15] += clock64();
    if (sw == 0 && lane < 4) s_uwin[b][lane] = w[lane];
    for (int c = 0; c < nCh; ++c) {
      const unsigned st = consumed % RING_PER_WARP;
      if (cyc) cyc[7] -= clock64();
      mbar_wait(smem_u32(&sh.ringFull[sw * RING_PER_WARP + st]), (consumed / RING_PER_WARP) & 1u);
      if (cyc) { const long long t = clock64(); cyc[7] += t; cyc[14] -= t; }
      const unsigned stage = ring + st * stageBytes;
      for (int r = 0; r < BY; r += 2) {            // two rows per batch (a stage row pitch inside a box is BX cells)
        const int row = (sw + c * NSW) * BY + r;
        if (row >= rowsMine) break;
        const int nr = min(min(2, BY - r), rowsMine - row);
        const unsigned rowS = stage + (unsigned)(r * BX * 8);
        unsigned* Urow = U + (size_t)row * UW;
        switch (nTW) {      // 64-cell groups per row: compile-time trip counts for the common window sizes
          case 4: pack_rows<4, 2>(rowS, BX * 8, (int)boxBytes, Urow, UW, nr, lane); break;
          case 5: pack_rows<5, 2>(rowS, BX * 8, (int)boxBytes, Urow, UW, nr, lane); break;
          case 6: pack_rows<6, 2>(rowS, BX * 8, (int)boxBytes, Urow, UW, nr, lane); break;
          case 8: pack_rows<8, 2>(rowS, BX * 8, (int)boxBytes, Urow, UW, nr, lane); break;
          case 9: pack_rows<9, 2>(rowS, BX * 8, (int)boxBytes, Urow, UW, nr, lane); break;
          case 10: pack_rows<10, 2>(rowS, BX * 8, (int)boxBytes, Urow, UW, nr, lane); break;
          default:      // any other width: one 64-cell group at a time
            for (int rr = 0; rr < nr; ++rr) {
              for (int t = 0; t < UW / 2; ++t) {
                unsigned w0 = 0u, w1 = 0u;
                if (t < nTW) {
                  const unsigned a = rowS + (unsigned)(rr * BX * 8 + (t / tPerBox) * (int)boxBytes + (t % tPerBox) * 512 + lane * 8);
                  const float2 lo = lds_f2(a), hi = lds_f2(a + 256u);
                  w0 = __ballot_sync(FULL, 2.f * lo.x > lo.y);
                  w1 = __ballot_sync(FULL, 2.f * hi.x > hi.y);
                }
                if (lane == 0) *reinterpret_cast<uint2*>(Urow + (size_t)rr * UW + 2 * t) = make_uint2(w0, w1);
              }
            }
        }
      }
      __syncwarp();
      if (cyc) cyc[14] += clock64();
      ++consumed;
      top_up();                                                   // refill the stage just drained
    }
    mbar_arrive(smem_u32(&sh.uFull[b]));                          // 32 * NSW arrivals: every lane's bitmap words are published
  }
}

// What the phases of one stage hand to each other (shared memory).  Every phase is an out-of-line function called
// from a thin driver: under the call ABI a callee only gets the registers its caller does not keep live across the
// call, and with the whole stage in one function the gather / blur loops were scheduled with a single load in flight.
struct StageCtx {
  double cx, cy, cth;      // centre of the search (proposal, or the coarse result)
  double xr0, yr0;         // window origin (:21-22)
  double thr;              // clamp threshold 0.5 * probMin (:44)
  int Wx, Wy, K0;          // field size, beams < maxRange
};

struct StageOut {
  double x, y, th, conf;
  int it, ia, ib;
};

// All static shared memory of the CTA behind one pointer: the per-particle driver keeps (almost) nothing live across
// the phase calls, so every phase gets the whole register file.
struct CtaShared {
  StreamShared ss;
  BlockScratch bs;
  StageCtx ctx;
  StageOut res[2];         // coarse / fine result of the current particle
  int uwin[2][4];          // union window (row0, col0, rows, cols) of the two bitmaps in flight
};

// ---- phases A-C: window geometry, index maps, occupied cells of the window
template <bool FAST, bool DENSE>
__device__ __noinline__ void window_phase(const MatchParams& P, CtaShared& sh, int stageId, int k, int& status) {
  const StageDev& S = P.st[stageId];
  StageCtx& ctx = sh.ctx;
  BlockScratch& bs = sh.bs;
  long long* cyc = P.dbgCycles ? P.dbgCycles + (size_t)blockIdx.x * 48 + 8 * stageId : nullptr;   // [0..15] phases of the two stages
  unsigned char* gslot = P.scratch + (size_t)blockIdx.x * P.slotBytes;
  const int tid = ctid(), lane = tid & 31, warp = tid >> 5;
  const double ul = S.unitLength;
  const int words = S.words, Pp = S.Ppitch;
  long long* sub = cyc ? cyc + (stageId == 0 ? 16 : 24) : nullptr;     // sub-phase slots (cyc is already offset by 8 for stage 1)
  (void)gslot; (void)lane; (void)warp; (void)ul; (void)words; (void)Pp; (void)sub;
  const double cx = ctx.cx, cy = ctx.cy;
  const unsigned* U = reinterpret_cast<const unsigned*>(gslot + P.gU) + (size_t)(k & 1) * P.URows * P.UW;   // this particle's union bitmap
  const int* uwin = sh.uwin[k & 1];
  // the fine stage is the last reader of the union bitmap (or the only stage of a field-only call)
  const unsigned uEmptyBar = (stageId == 1 || P.fieldOnlyStage == 0) ? smem_u32(&sh.ss.uEmpty[k & 1]) : 0u;
  // ---- A. geometry of the search window (ScanMatcher_OGBased.py:21-28)
  const double xr0 = dsub(cx, P.R), xr1 = dadd(cx, P.R);
  const double yr0 = dsub(cy, P.R), yr1 = dadd(cy, P.R);
  const int Wx = (int)ddiv(dsub(xr1, xr0), ul) + 1;
  const int Wy = (int)ddiv(dsub(yr1, yr0), ul) + 1;
  if (tid == 0) { ctx.xr0 = xr0; ctx.yr0 = yr0; ctx.Wx = Wx; ctx.Wy = Wy; }     // published by the barriers below
  int mx0 = (int)rint(ddiv(dsub(xr0, P.mapX0), P.unit)), mx1 = (int)rint(ddiv(dsub(xr1, P.mapX0), P.unit));
  int my0 = (int)rint(ddiv(dsub(yr0, P.mapY0), P.unit)), my1 = (int)rint(ddiv(dsub(yr1, P.mapY0), P.unit));
  if (xr0 < P.mapX0 || xr1 > P.mapX1 || yr0 < P.mapY0 || yr1 > P.mapY1) status |= SLAM_ST_WINDOW_OUTSIDE_MAP;
  mx0 = max(mx0, 0); my0 = max(my0, 0);
  mx1 = min(mx1, P.G); my1 = min(my1, P.G);
  mx1 = min(mx1, mx0 + S.Wmap); my1 = min(my1, my0 + S.Wmap);
  const int ncols = max(mx1 - mx0, 0), nrows = max(my1 - my0, 0);

  unsigned* bits = buf<FAST, unsigned>(S.oBits, gslot, S.gBits);
  unsigned* bitsT = buf<FAST, unsigned>(S.oBitsT, gslot, S.gBitsT);    // [col][WT] transposed
  unsigned* dil = buf<FAST, unsigned>(S.oDil, gslot, S.gDil);         // activity bitmap
  const int WT = S.WT;
  short* rowMap = sbuf<short>(S.oRow);
  short* colMap = sbuf<short>(S.oCol);
  double* Pf = DENSE ? sbuf<double>(S.oP) : reinterpret_cast<double*>(gslot + S.gP);
  constexpr bool dense = DENSE;                      // coarse: dense field in shared memory, ungated gathers

  csync();  // previous users of the arena are done
  if (cyc && tid == 0) cyc[0] -= clock64();
  SubCyc sc;
  sc.start(sub);

  // ---- B. clear bitmaps, float64 index maps (:36 via :173-176) -- rows of OccupancyGridX are identical -- and the
  //         statistics of the maps that select the scatter path
  int* stat = bs.ibcast;            // [0] min, [1] max of colMap[j] - j; [2], [3] same for rows; [4] "not monotone"
  if (tid == 0) { stat[0] = 1 << 30; stat[1] = -(1 << 30); stat[2] = 1 << 30; stat[3] = -(1 << 30); stat[4] = 0; }
  for (int i = tid; i < Wy * words; i += NTC) bits[i] = 0u;
  for (int i = tid; i < Wx * WT; i += NTC) bitsT[i] = 0u;
  if (dense)
    for (int i = tid; i < Wy * Pp; i += NTC) Pf[i] = S.B2;
  csync();
  {
    int mn = 1 << 30, mx = -(1 << 30);
    for (int j = tid; j < ncols; j += NTC) {
      int c = (int)ddiv(dsub(P.gridX[mx0 + j], xr0), ul);
      if (c < 0 || c >= Wx) { status |= SLAM_ST_INDEX_OUT_OF_FIELD; c = min(max(c, 0), Wx - 1); }
      colMap[j] = (short)c;
      mn = min(mn, c - j); mx = max(mx, c - j);
    }
    mn = __reduce_min_sync(FULL, mn); mx = __reduce_max_sync(FULL, mx);
    if (lane == 0 && mn <= mx) { atomicMin(&stat[0], mn); atomicMax(&stat[1], mx); }
    mn = 1 << 30; mx = -(1 << 30);
    for (int i = tid; i < nrows; i += NTC) {
      int c = (int)ddiv(dsub(P.gridY[my0 + i], yr0), ul);
      if (c < 0 || c >= Wy) { status |= SLAM_ST_INDEX_OUT_OF_FIELD; c = min(max(c, 0), Wy - 1); }
      rowMap[i] = (short)c;
      mn = min(mn, c - i); mx = max(mx, c - i);
    }
    mn = __reduce_min_sync(FULL, mn); mx = __reduce_max_sync(FULL, mx);
    if (lane == 0 && mn <= mx) { atomicMin(&stat[2], mn); atomicMax(&stat[3], mx); }
  }
  csync();
  sc.mark(0);      // B clear + maps

  // ---- C. occupied cells of this stage's window (:29-37) out of the particle's union bitmap (plain layout: word w of
  //         a row = cells 32w .. 32w+31), which the stream warp built in the background.  Three paths, same result:
  //   shift   both index maps are "j + D or j + D - 1" (fine stage: map and field share the lattice, truncation
  //           noise decides): whole words move -- masked funnel shifts, OR of the <= 2 source rows; no atomics
  //   range   monotone many-to-one maps (coarse stage): OR of the source-row range, then one bit-range test per cell
  //   generic one atomicOr per occupied cell (any map)
  {
    const int uy0 = uwin[0], ux0 = uwin[1], unr = uwin[2], unc = uwin[3];
    if (my0 < uy0 || my1 > uy0 + unr || mx0 < ux0 || mx1 > ux0 + unc) {
      if (nrows > 0 && ncols > 0) status |= SLAM_ST_WINDOW_OUTSIDE_MAP;     // union window was clipped
    }
    const int UW = P.UW;
    const int ax = mx0 - ux0, ay = my0 - uy0;          // window origin inside the union bitmap
    const int Dc = stat[1], Dr = stat[3];
    int mode = 0;
    if (ncols > 0 && nrows > 0 && ax >= 0 && ay >= 0 && !P.forceGenericScatter) {
      if (UW <= 32 && words <= 32 && Dc - stat[0] <= 1 && Dr - stat[2] <= 1) mode = 1;
      else if (S.auxOK && UW <= 64) mode = 2;
    }
    short* colLo = sbuf<short>(S.oAux);
    short* colHi = colLo + S.auxPitch;
    short* rowLo = colHi + S.auxPitch;
    short* rowHi = rowLo + S.auxPitch;
    unsigned* rowbuf = reinterpret_cast<unsigned*>(rowHi + S.auxPitch);     // [Wy][UW]
    if (mode == 2) {      // source ranges of every field column / row; falls back if a map is not monotone
      for (int i = tid; i < S.auxPitch; i += NTC) { colLo[i] = 0; colHi[i] = 0; rowLo[i] = 0; rowHi[i] = 0; }
      csync();
      int bad = 0;
      for (int j = tid; j < ncols; j += NTC) {
        const int c = colMap[j], pv = j > 0 ? (int)colMap[j - 1] : -1, nx = j + 1 < ncols ? (int)colMap[j + 1] : 1 << 20;
        if (c < pv) bad = 1;
        if (c != pv) colLo[c] = (short)j;
        if (c != nx) colHi[c] = (short)(j + 1);
      }
      for (int i = tid; i < nrows; i += NTC) {
        const int c = rowMap[i], pv = i > 0 ? (int)rowMap[i - 1] : -1, nx = i + 1 < nrows ? (int)rowMap[i + 1] : 1 << 20;
        if (c < pv) bad = 1;
        if (c != pv) rowLo[c] = (short)i;
        if (c != nx) rowHi[c] = (short)(i + 1);
      }
      csync();
      for (int c = tid; c < Wx; c += NTC)
        if (colHi[c] - colLo[c] > 32) bad = 1;
      for (int c = tid; c < Wy; c += NTC)
        if (rowHi[c] - rowLo[c] > 32) bad = 1;
      if (bad) stat[4] = 1;
      csync();
      if (stat[4]) mode = 0;
    }
    if (mode == 1) {
      // masks: bit k of word t <=> union column 32t + k is window column j = 32t + k - ax with colMap[j] - j == Dc (A) / Dc - 1 (B)
      for (int t = warp; t < UW; t += NWC) {
        const int j = 32 * t + lane - ax;
        const bool in = j >= 0 && j < ncols;
        const int off = in ? (int)colMap[j] - j : 0;
        const unsigned mA = __ballot_sync(FULL, in && off == Dc), mB = __ballot_sync(FULL, in && off == Dc - 1);
        if (lane == 0) { bs.maskA[t] = mA; bs.maskB[t] = mB; }
      }
      csync();
      const int sA = Dc - ax;                          // field column = union column + sA (A) / + sA - 1 (B)
      const unsigned mA = lane < UW ? bs.maskA[lane] : 0u, mB = lane < UW ? bs.maskB[lane] : 0u;
      constexpr int RU4 = 12;                          // field rows per warp iteration: independent L2 load chains
      for (int fr0 = warp * RU4; fr0 < Wy; fr0 += NWC * RU4) {
        unsigned x[RU4];
#pragma unroll
        for (int q = 0; q < RU4; ++q) {
          const int fr = fr0 + q;
          const int i1 = fr - Dr, i2 = i1 + 1;         // the <= 2 window rows that map onto field row fr
          const bool v1 = fr < Wy && i1 >= 0 && i1 < nrows && rowMap[i1] == fr && i1 + ay < unr;
          const bool v2 = fr < Wy && i2 >= 0 && i2 < nrows && rowMap[i2] == fr && i2 + ay < unr;
          unsigned xa = 0u, xb = 0u;
          if (lane < UW) {
            if (v1) xa = __ldcg(U + (size_t)(i1 + ay) * UW + lane);
            if (v2) xb = __ldcg(U + (size_t)(i2 + ay) * UW + lane);
          }
          x[q] = xa | xb;
        }
        const int baseA = 32 * lane - sA, baseB = baseA + 1;
        const int tA = baseA >> 5, tB = baseB >> 5;
        const bool okA0 = tA >= 0 && tA <= 31, okA1 = tA + 1 >= 0 && tA + 1 <= 31;
        const bool okB0 = tB >= 0 && tB <= 31, okB1 = tB + 1 >= 0 && tB + 1 <= 31;
#pragma unroll
        for (int q = 0; q < RU4; ++q) {
          const unsigned a = x[q] & mA, b = x[q] & mB;
          unsigned loA = __shfl_sync(FULL, a, tA & 31), hiA = __shfl_sync(FULL, a, (tA + 1) & 31);
          unsigned loB = __shfl_sync(FULL, b, tB & 31), hiB = __shfl_sync(FULL, b, (tB + 1) & 31);
          if (!okA0) loA = 0u;
          if (!okA1) hiA = 0u;
          if (!okB0) loB = 0u;
          if (!okB1) hiB = 0u;
          const unsigned out = __funnelshift_r(loA, hiA, baseA & 31) | __funnelshift_r(loB, hiB, baseB & 31);
          if (lane < words && fr0 + q < Wy) bits[(fr0 + q) * words + lane] = out;
        }
      }
    } else if (mode == 2) {
      for (int q = tid; q < Wy * UW; q += NTC) {       // OR of the window rows that map onto each field row
        const int fr = q / UW, t = q - fr * UW;
        unsigned x = 0u;
        const int hiR = min((int)rowHi[fr], unr - ay);
        for (int i0 = rowLo[fr]; i0 < hiR; i0 += 8) {
          unsigned y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = i0 + e < hiR ? __ldcg(U + (size_t)(i0 + e + ay) * UW + t) : 0u;
#pragma unroll
          for (int e = 0; e < 8; ++e) x |= y[e];
        }
        rowbuf[q] = x;
      }
      csync();
      for (int q0 = warp * 32; q0 < Wy * words * 32; q0 += NWC * 32) {     // one field cell per lane, 32 cells = 1 word
        const int wq = q0 >> 5, fr = wq / words, g = wq - fr * words;
        const int c = 32 * g + lane;
        bool on = false;
        if (c < Wx) {
          const int lo = colLo[c] + ax, len = colHi[c] - colLo[c];
          if (len > 0) {
            const int t0 = lo >> 5;
            const unsigned w0 = rowbuf[fr * UW + t0], w1 = t0 + 1 < UW ? rowbuf[fr * UW + t0 + 1] : 0u;
            const unsigned v = __funnelshift_r(w0, w1, lo & 31);
            on = (v & (len >= 32 ? 0xffffffffu : ((1u << len) - 1u))) != 0u;
          }
        }
        const unsigned wv = __ballot_sync(FULL, on);
        if (lane == 0) bits[fr * words + g] = wv;
      }
    } else {
      const int wlo = max(ax, 0) >> 5, whi = min(((mx1 - 1 - ux0) >> 5) + 1, UW);
      const int nw = max(whi - wlo, 0);
      const int rlo = max(ay, 0), rhi = min(my1 - uy0, unr);
      const int total = max(rhi - rlo, 0) * nw;
      // a warp takes 32 bitmap words at a time and spreads their set bits evenly over its lanes: lane j handles the
      // j-th set bit of the group (owner word by binary search over the popcount prefix, bit by __fns)
      auto loadw = [&](int base) -> unsigned {
        const int idx = base + lane;
        if (idx >= total) return 0u;
        const int rr = idx / nw;
        return __ldcg(U + (size_t)(rlo + rr) * UW + wlo + (idx - rr * nw));
      };
      unsigned wnext = loadw(warp * 32);
      for (int base = warp * 32; base < total; base += NWC * 32) {
        const unsigned wmine = wnext;
        wnext = loadw(base + NWC * 32);                    // next group's words are in flight while this one is expanded
        const int cnt = __popc(wmine);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(FULL, incl, d);
          if (lane >= d) incl += t;
        }
        const int tot = __shfl_sync(FULL, incl, 31);
        const int excl = incl - cnt;
        for (int j0 = 0; j0 < tot; j0 += 32) {
          const int j = min(j0 + lane, tot - 1);
          int L = 0;
#pragma unroll
          for (int step = 16; step >= 1; step >>= 1) {
            const int probe = __shfl_sync(FULL, incl, L + step - 1);
            if (probe <= j) L += step;
          }
          const unsigned wv = __shfl_sync(FULL, wmine, L);
          const int kth = j - __shfl_sync(FULL, excl, L);
          if (j0 + lane < tot) {
            const int bit = __fns(wv, 0, kth + 1);
            const int id2 = base + L;
            const int rr = id2 / nw, wi = wlo + (id2 - rr * nw);
            const int i = uy0 + rlo + rr - my0;                               // window row
            const int jc = 32 * wi + bit - ax;                                // window column
            if (jc >= 0 && jc < ncols) {
              const int fr = (int)rowMap[i];
              const int c = colMap[jc];
              atomicOr(&bits[fr * words + (c >> 5)], 1u << (c & 31));
              atomicOr(&bitsT[c * WT + (fr >> 5)], 1u << (fr & 31));
            }
          }
        }
      }
    }
    if (mode != 0) {      // transposed bitmap by 32x32 block transposes of the row-major one (empty blocks skipped)
      csync();
      sc.mark(1);    // C scatter
      const int nrb = (Wy + 31) >> 5, nwb = (Wx + 31) >> 5;
      for (int q = warp; q < nrb * nwb; q += NWC) {
        const int rb = q / nwb, wb = q - rb * nwb;
        const int row = 32 * rb + lane;
        const unsigned w = row < Wy ? bits[row * words + wb] : 0u;
        if (__ballot_sync(FULL, w != 0u) == 0u) continue;          // bitsT was cleared in B
        // 32x32 bit-matrix transpose across the warp: five butterfly exchanges (lane k <-> k ^ j swap sub-blocks)
        unsigned out = w;
#pragma unroll
        for (int j = 16, m = 0x0000ffff; j; j >>= 1, m ^= (m << j)) {
          const unsigned y = __shfl_xor_sync(FULL, out, j);
          out = (lane & j) ? (((y >> j) & (unsigned)m) | (out & ~(unsigned)m)) : ((out & (unsigned)m) | ((y << j) & ~(unsigned)m));
        }
        const int col = 32 * wb + lane;
        if (col < Wx) bitsT[col * WT + rb] = out;
      }
    }
  }
  csync();
  sc.mark(2);      // C transposes (or the whole generic scatter)
  if (uEmptyBar && tid == 0) mbar_arrive(uEmptyBar);   // last reader of the union bitmap: the stream warp may reuse it
  if (cyc && tid == 0) { long long t = clock64(); cyc[0] += t; cyc[1] -= t; }
}

// ---- phases D-F: blur, min / clamp, beam end points
template <bool FAST, bool DENSE>
__device__ __noinline__ void field_phase(const MatchParams& P, CtaShared& sh, int stageId, int p, int& status) {
  const StageDev& S = P.st[stageId];
  StageCtx& ctx = sh.ctx;
  BlockScratch& bs = sh.bs;
  long long* cyc = P.dbgCycles ? P.dbgCycles + (size_t)blockIdx.x * 48 + 8 * stageId : nullptr;   // [0..15] phases of the two stages
  unsigned char* gslot = P.scratch + (size_t)blockIdx.x * P.slotBytes;
  const int tid = ctid(), lane = tid & 31, warp = tid >> 5;
  const double ul = S.unitLength;
  const int words = S.words, Pp = S.Ppitch;
  long long* sub = cyc ? cyc + (stageId == 0 ? 16 : 24) : nullptr;     // sub-phase slots (cyc is already offset by 8 for stage 1)
  (void)gslot; (void)lane; (void)warp; (void)ul; (void)words; (void)Pp; (void)sub;
  const double cx = ctx.cx, cy = ctx.cy, cth = ctx.cth;
  const int Wx = ctx.Wx, Wy = ctx.Wy, r = S.r;
  unsigned* dil = buf<FAST, unsigned>(S.oDil, gslot, S.gDil);         // activity bitmap
  double* Pf = DENSE ? sbuf<double>(S.oP) : reinterpret_cast<double*>(gslot + S.gP);
  constexpr bool dense = DENSE;
  // ---- D. separable blur in scipy's order, only where the result can differ from the background.
  double mn = 0.0;                 // every field value is <= 0
  int nActive = 0;
  double thr = 0.0;
  {
    int* counter = &bs.ibcast[4];
    switch (r) {
      case 2: blur_stage<2, FAST, DENSE>(S, gslot, Pp, Wx, Wy, counter, mn, nActive, thr, sub); break;
      case 4: blur_stage<4, FAST, DENSE>(S, gslot, Pp, Wx, Wy, counter, mn, nActive, thr, sub); break;
      case 8: blur_stage<8, FAST, DENSE>(S, gslot, Pp, Wx, Wy, counter, mn, nActive, thr, sub); break;
      default: blur_stage<0, FAST, DENSE>(S, gslot, Pp, Wx, Wy, counter, mn, nActive, thr, sub); break;
    }
  }
  // ---- E. probMin, clamp (:43-44).  Normally done inside the blur (probMin == B2); only a window without a single
  //         background cell needs the explicit minimum + a second pass.
  {
    const double probMin = block_min(mn, bs);          // also the barrier that publishes the field
    if (probMin < S.B2) status |= SLAM_ST_INDEX_OUT_OF_FIELD;   // cannot happen (monotone rounding); fail loudly
    if (thr == 0.0) {
      thr = dmul(0.5, probMin);
      for (int i = tid; i < Wy * Wx; i += NTC) {
        const int yy = i / Wx, xx = i - yy * Wx;
        if ((dil[yy * words + (xx >> 5)] >> (xx & 31)) & 1u) {
          double* q = Pf + (size_t)yy * Pp + xx;
          if (*q > thr) *q = 0.0;
        }
      }
      csync();
    }
  }
  if (cyc && tid == 0) { long long t = clock64(); cyc[1] += t; cyc[2] -= t; }
  if (P.dbgProb[stageId]) {
    double* d = P.dbgProb[stageId] + (size_t)p * S.Wmax * S.Wmax;
    for (int i = tid; i < Wy * Wx; i += NTC) {
      const int yy = i / Wx, xx = i - yy * Wx;
      const bool on = dense || ((dil[yy * words + (xx >> 5)] >> (xx & 31)) & 1u);
      d[yy * S.Wmax + xx] = on ? Pf[(size_t)yy * Pp + xx] : S.B2;
    }
    if (tid == 0) { P.dbgDims[stageId][2 * p] = Wy; P.dbgDims[stageId][2 * p + 1] = Wx; }
  }

  // ---- F. beam end points (:81-89) with order-preserving compaction of beams < maxRange
  double* dxs = sbuf<double>(S.oDx);
  double* dys = sbuf<double>(S.oDy);
  int K0;
  {
    const int K = P.fieldOnlyStage >= 0 ? 0 : P.K;      // slam_field_build has no scan
    const double start = dsub(cth, P.fovHalf), stop = dadd(cth, P.fovHalf);
    const double step = ddiv(dsub(stop, start), (double)(K - 1));
    int base = 0;
    for (int k0 = 0; k0 < K; k0 += NTC) {     // K <= 512 -> one trip
      const int k = k0 + tid;
      bool keep = false;
      double ddx = 0, ddy = 0;
      if (k < K) {
        const double rm = P.ranges[k];
        keep = rm < P.maxRange;
        const double ang = (k == K - 1) ? stop : dadd(dmul((double)k, step), start);   // np.linspace
        double sn, cs;
        sincos(ang, &sn, &cs);
        const double px = dadd(cx, dmul(cs, rm)), py = dadd(cy, dmul(sn, rm));
        ddx = dsub(px, cx); ddy = dsub(py, cy);                                          // (px - ox) of :169
      }
      unsigned bal = __ballot_sync(FULL, keep);
      csync();
      if (lane == 0) bs.ival[warp] = __popc(bal);
      csync();
      int before = base;
      for (int w2 = 0; w2 < warp; ++w2) before += bs.ival[w2];
      int tot = 0;
      for (int w2 = 0; w2 < NWC; ++w2) tot += bs.ival[w2];
      if (keep) {
        int pos = before + __popc(bal & ((1u << lane) - 1u));
        dxs[pos] = ddx; dys[pos] = ddy;
      }
      base += tot;
    }
    K0 = base;
  }
  if (tid == 0) { ctx.K0 = K0; ctx.thr = thr; }
  csync();

  if (cyc && tid == 0) cyc[2] += clock64();}

// ---- phase G: per-theta point lists and the score volume
template <bool FAST, bool DENSE>
__device__ __noinline__ void correlate_phase(const MatchParams& P, CtaShared& sh, int stageId, int p, int& status) {
  const StageDev& S = P.st[stageId];
  StageCtx& ctx = sh.ctx;
  BlockScratch& bs = sh.bs;
  long long* cyc = P.dbgCycles ? P.dbgCycles + (size_t)blockIdx.x * 48 + 8 * stageId : nullptr;   // [0..15] phases of the two stages
  unsigned char* gslot = P.scratch + (size_t)blockIdx.x * P.slotBytes;
  const int tid = ctid(), lane = tid & 31, warp = tid >> 5;
  const double ul = S.unitLength;
  const int words = S.words, Pp = S.Ppitch;
  long long* sub = cyc ? cyc + (stageId == 0 ? 16 : 24) : nullptr;     // sub-phase slots (cyc is already offset by 8 for stage 1)
  (void)gslot; (void)lane; (void)warp; (void)ul; (void)words; (void)Pp; (void)sub;
  const double cx = ctx.cx, cy = ctx.cy, xr0 = ctx.xr0, yr0 = ctx.yr0, thr = ctx.thr;
  const int Wx = ctx.Wx, Wy = ctx.Wy, K0 = ctx.K0;
  // ---- G. per-theta lists + score volume, TB thetas at a time
  double* scores = S.needScores ? buf<FAST, double>(S.oScores, gslot, S.gScores) : nullptr;
  double* dvol = P.dbgVol[stageId] ? P.dbgVol[stageId] + (size_t)p * S.nPoses : nullptr;
  const int nOff = S.nOff, nOff2 = nOff * nOff;
  const double* rv = (stageId == 0) ? P.rv : nullptr;
  const double* tw = (stageId == 0 && P.tw) ? P.tw + (size_t)p * nOff2 : nullptr;
  if (lane == 0) { bs.wbest[warp] = 0.0; bs.wbestIdx[warp] = -1; }      // own slot: ordered by program order within the warp
  if (tid == 0) bs.nanFlag = 0;                                          // read after the barriers below
  ScoreArgs SA;
  SA.bs = &bs;
  SA.ptsSlot = P.dbgCycles ? P.dbgCycles + (size_t)blockIdx.x * 48 + 44 : nullptr;
  SA.gslot = gslot; SA.rv = rv; SA.tw = tw; SA.dvol = dvol; SA.B2 = S.B2; SA.thr = thr;
  SA.gP = S.gP; SA.gDil = S.gDil; SA.gScores = S.gScores;
  SA.oLists = S.oLists; SA.oCnt = S.oCnt; SA.oP = S.oP; SA.oDil = S.oDil; SA.oScores = S.oScores;
  SA.needScores = S.needScores;
  SA.Kpad = S.Kpad; SA.Pp = Pp; SA.words = words; SA.nHalf = S.nHalf; SA.nOff = nOff; SA.nGrpPad = S.nGrpPad;
  // exact branch-and-bound only where nothing but the argmax is needed (fine stage, no volume dump)
  const bool prune = !DENSE && stageId == 1 && !S.needScores && !dvol && !rv && !tw && !P.noPrune;
  SA.bestP = &bs.incumbent; SA.bestS = smem_u32(&bs.incumbent);
  if (tid == 0) bs.incumbent = -INFINITY;
  ListArgs LA;
  LA.cosT = S.cosT; LA.sinT = S.sinT;
  LA.ox = cx; LA.oy = cy; LA.bx = xr0; LA.by = yr0; LA.ul = ul;
  LA.oDx = S.oDx; LA.oDy = S.oDy; LA.oLists = S.oLists; LA.oCnt = S.oCnt;
  LA.K0 = K0; LA.nHalf = S.nHalf; LA.Wx = Wx; LA.Wy = Wy; LA.Kpad = S.Kpad;
  for (int t0 = 0; t0 < S.nTheta; t0 += S.TB) {
    const int nt = min(S.TB, S.nTheta - t0);
    if (cyc && tid == 0) cyc[3] -= clock64();
    LA.nt = nt; LA.t0 = t0;
    if (S.E == 8) lists_batch<8>(LA, status);
    else lists_batch<16>(LA, status);
    if (tid == 0) bs.taskCounter = 0;          // chunk counter of the score phase, published by the barrier below
    csync();
    if (cyc && tid == 0) { long long t = clock64(); cyc[3] += t; cyc[4] -= t; }
    SA.nt = nt; SA.t0 = t0;
    if (prune) score_batch<FAST, DENSE, true>(SA);
    else score_batch<FAST, DENSE, false>(SA);
    csync();
    if (cyc && tid == 0) cyc[4] += clock64();
  }
}

// ---- phase H: argmax / softmax-CDF sample, confidence, matched pose
template <bool FAST, bool DENSE>
__device__ __noinline__ void select_phase(const MatchParams& P, CtaShared& sh, int stageId, int p, int& status) {
  const StageDev& S = P.st[stageId];
  StageCtx& ctx = sh.ctx;
  BlockScratch& bs = sh.bs;
  long long* cyc = P.dbgCycles ? P.dbgCycles + (size_t)blockIdx.x * 48 + 8 * stageId : nullptr;   // [0..15] phases of the two stages
  unsigned char* gslot = P.scratch + (size_t)blockIdx.x * P.slotBytes;
  const int tid = ctid(), lane = tid & 31, warp = tid >> 5;
  const double ul = S.unitLength;
  const int words = S.words, Pp = S.Ppitch;
  long long* sub = cyc ? cyc + (stageId == 0 ? 16 : 24) : nullptr;     // sub-phase slots (cyc is already offset by 8 for stage 1)
  (void)gslot; (void)lane; (void)warp; (void)ul; (void)words; (void)Pp; (void)sub;
  const double cx = ctx.cx, cy = ctx.cy, cth = ctx.cth;
  const int nOff = S.nOff, nOff2 = nOff * nOff;
  double* scores = S.needScores ? buf<FAST, double>(S.oScores, gslot, S.gScores) : nullptr;
  const bool sample = stageId == 0 && P.uniforms != nullptr;      // the fine stage always takes the argmax (:73)
  const double uniform = sample ? P.uniforms[p] : 0.0;
  StageOut out;
  if (cyc && tid == 0) cyc[5] -= clock64();
  SubCyc sc;
  sc.start(sub);
  // ---- H. select (:133-141)
  // NaN scores: np.random.choice raises ValueError (:138); np.argmax just returns the first NaN (:134), so only the
  // sampled stage reports them (the correlate phase ended with a barrier)
  if (bs.nanFlag && sample) status |= SLAM_ST_NAN_SCORE;
  int chosen;
  {  // first maximum in C order: the per-warp maxima of the correlate phase
    double b = bs.wbest[0];
    int bi = bs.wbestIdx[0];
    for (int w2 = 1; w2 < NWC; ++w2) {
      double ob = bs.wbest[w2];
      int oi = bs.wbestIdx[w2];
      if (first_max_better(ob, oi, b, bi)) { b = ob; bi = oi; }
    }
    chosen = bi;
    csync();
  }
  sc.mark(8);      // H argmax
  double conf = 0.0;
  if (S.needScores) {
    const int n = S.nPoses;
    for (int i = tid; i < n; i += NTC) scores[i] = exp(scores[i]);
    csync();
    sc.mark(9);    // H exp
    double* leafSum = sbuf<double>(S.oLeaf);
    // leaves of numpy's pairwise sum: eight threads per leaf, thread l owns the running lane r[l]
    for (int g0 = warp * 4; g0 < S.nLeaves; g0 += NWC * 4) {
      const int g = g0 + (lane >> 3), l = lane & 7;
      const bool valid = g < S.nLeaves;
      const int2 lf = valid ? S.leaves[g] : make_int2(0, 0);
      const int off = lf.x, n = lf.y, m = n - (n & 7);
      double r = (n >= 8) ? scores[off + l] : 0.0;
      for (int i = 8; i < m; i += 8) r = dadd(r, scores[off + i + l]);
      r = dadd(r, __shfl_xor_sync(FULL, r, 1));        // (r0+r1), (r2+r3), ...
      r = dadd(r, __shfl_xor_sync(FULL, r, 2));        // ((r0+r1)+(r2+r3)), ...
      r = dadd(r, __shfl_xor_sync(FULL, r, 4));
      if (n < 8) r = 0.0;
      for (int i = (n < 8 ? 0 : m); i < n; ++i) r = dadd(r, scores[off + i]);     // n < 8: plain sequential sum; else the tail
      if (valid && l == 0) leafSum[g] = r;
    }
    // operands of the combine tree (global tables): fetched by warp 0 before the barrier, not level by level after it
    int2 opA = make_int2(0, 0), opB = make_int2(0, 0);
    const bool opsInRegs = S.nOps <= 64;
    if (warp == 0 && opsInRegs) {
      if (lane < S.nOps) opA = S.ops[lane];
      if (lane + 32 < S.nOps) opB = S.ops[lane + 32];
    }
    csync();
    if (warp == 0) {   // combine the leaves along numpy's recursion tree, one tree level at a time
      for (int l = 0; l < S.nLevels; ++l) {
        const int o0 = S.levelStart[l], o1 = S.levelStart[l + 1];
        if (opsInRegs) {
          if (lane >= o0 && lane < o1) leafSum[S.nLeaves + lane] = dadd(leafSum[opA.x], leafSum[opA.y]);
          if (lane + 32 >= o0 && lane + 32 < o1) leafSum[S.nLeaves + lane + 32] = dadd(leafSum[opB.x], leafSum[opB.y]);
        } else {
          for (int o = o0 + lane; o < o1; o += 32) {
            const int2 op = S.ops[o];
            leafSum[S.nLeaves + o] = dadd(leafSum[op.x], leafSum[op.y]);
          }
        }
        __syncwarp();
      }
      if (lane == 0) bs.bcast[0] = leafSum[S.nOps ? S.nLeaves + S.nOps - 1 : 0];     // root (a single leaf: no ops)
    }
    csync();
    conf = bs.bcast[0];
    sc.mark(10);   // H pairwise sum of exp
    if (sample) {
      // np.random.choice: p = e / e.sum(); cdf = cumsum(p) (sequential); cdf /= cdf[-1]; searchsorted(u, 'right').
      // Located with a parallel prefix and certified by a margin: with p >= 0 both the sequential cumsum and the tree
      // prefix lie within (n - 1) * 2^-53 * sum(p) of the exact prefix (Higham, Accuracy and Stability, 4.2) and the
      // normalisation adds 2 ulp, so the two normalised CDFs differ by < (2n + 4) * 2^-53 (< 5e-12 for n <= 21870).
      // If u is farther than tol = max(1e-10, 16 n 2^-53) from the CDF values on both sides of the located step, the
      // exact monotone CDF crosses u at the same index; otherwise thread 0 walks the exact sequential definition.
      for (int i = tid; i < n; i += NTC) scores[i] = ddiv(scores[i], conf);      // p = e / e.sum()
      csync();
      const int chunk = (n + NTC - 1) / NTC;
      const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
      double part = 0.0;
      for (int i = lo; i < hi; ++i) part += scores[i];
      // block exclusive scan of the chunk sums
      double incl = part;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        double v = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += v;
      }
      if (lane == 31) bs.dval[warp] = incl;
      if (tid == 0) { bs.ibcast[0] = -1; bs.ibcast[1] = 0; }
      csync();
      double wbase = 0.0, total = 0.0;
      for (int w2 = 0; w2 < NWC; ++w2) {
        if (w2 < warp) wbase += bs.dval[w2];
        total += bs.dval[w2];
      }
      const double before = wbase + incl - part;
      const double tol = fmax(1e-10, 16.0 * (double)n * 1.1102230246251565e-16);
      const double ut = uniform * total;
      if (hi > lo && before <= ut && (before + part > ut || hi == n)) {
        double c = before, prev = before;
        int found = -1;
        for (int i = lo; i < hi; ++i) {
          prev = c;
          c += scores[i];
          if (c > ut) { found = i; break; }
        }
        if (found >= 0) {
          bool amb = (uniform - prev / total) < tol || (c / total - uniform) < tol;
          // several threads can only claim when chunks tie within rounding -> treated as ambiguous below
          int old = atomicExch(&bs.ibcast[0], found);
          if (old != -1 || amb) bs.ibcast[1] = 1;
        }
      }
      csync();
      int idx = bs.ibcast[0];
      bool exact = P.forceExactCdf || idx < 0 || bs.ibcast[1];
      csync();
      if (exact) {
        if (tid == 0) {
          double c = 0.0;
          for (int i = 0; i < n; ++i) c = dadd(c, scores[i]);
          const double last = c;
          c = 0.0;
          int found = n;
          for (int i = 0; i < n; ++i) {
            c = dadd(c, scores[i]);
            if (ddiv(c, last) > uniform) { found = i; break; }
          }
          bs.ibcast[0] = min(found, n - 1);
        }
        csync();
        idx = bs.ibcast[0];
      }
      chosen = idx;
      sc.mark(11);   // H CDF inversion
    }
  }
  if (chosen < 0) chosen = 0;
  const int it = chosen / nOff2, rem = chosen - it * nOff2;
  const int ia = rem / nOff, ib = rem - ia * nOff;
  out.it = it; out.ia = ia; out.ib = ib;
  out.x = dadd(cx, dmul((double)(ib - S.nHalf), ul));      // :142-143
  out.y = dadd(cy, dmul((double)(ia - S.nHalf), ul));
  out.th = dadd(cth, S.thetas[it]);
  out.conf = conf;
  if (tid == 0) sh.res[stageId] = out;      // read after the next barrier
  if (cyc && tid == 0) cyc[5] += clock64();
}

template <bool FAST, bool DENSE>
__device__ __forceinline__ void run_stage(const MatchParams& P, CtaShared& sh, int stageId, int p, int k, int& status) {
  csync();                // the previous stage's readers of ctx are done, its result is published
  if (ctid() == 0) {
    StageCtx& ctx = sh.ctx;
    if (stageId == 0 || P.fieldOnlyStage >= 0) { ctx.cx = P.estPose[3 * p]; ctx.cy = P.estPose[3 * p + 1]; ctx.cth = P.estPose[3 * p + 2]; }
    else { ctx.cx = sh.res[0].x; ctx.cy = sh.res[0].y; ctx.cth = sh.res[0].th; }     // centred on the coarse result (:66-73)
  }
  csync();
  window_phase<FAST, DENSE>(P, sh, stageId, k, status);
  field_phase<FAST, DENSE>(P, sh, stageId, p, status);
  if (P.fieldOnlyStage >= 0) return;      // slam_field_build: the field was dumped by field_phase
  correlate_phase<FAST, DENSE>(P, sh, stageId, p, status);
  select_phase<FAST, DENSE>(P, sh, stageId, p, status);
}

// Cold start: nothing overlaps the stream of a CTA's FIRST particle, so the compute warps pack the tail rows of its
// union window themselves, 2 rows per warp at a time.
__device__ __noinline__ void cold_start(const MatchParams& P) {
  unsigned char* gslot = P.scratch + (size_t)blockIdx.x * P.slotBytes;
  unsigned* Ubuf0 = reinterpret_cast<unsigned*>(gslot + P.gU);
  long long* cyc = P.dbgCycles ? P.dbgCycles + (size_t)blockIdx.x * 48 : nullptr;
  {
    const int p = blockIdx.x, lane = ctid() & 31, warp = ctid() >> 5;
    if (cyc && ctid() == 0) { const long long t = clock64(); cyc[6] -= t; cyc[29] -= t; }     // accounted as waiting for the union bitmap
    int w[4];
    union_window(P, p, w);
    const int r0 = first_rows_by_stream(w[2], P.ringRows);
    const int nTW = P.UWcells / 64, UW = P.UW, nc = w[3];
    const size_t lat = P.slots ? P.slots[p] : p;
    const float2* base = reinterpret_cast<const float2*>(P.grid) + (lat * P.G + w[0]) * P.pitch + w[1];
    if (nTW <= 10) {
      // software-pipelined: the 16-byte loads of a warp's NEXT row are in flight while it thresholds and packs this one
      // (measured: compute-side pack 187 k -> 130 k cycles per CTA)
      auto load_row = [&](int row, float4 (&v)[10]) {
        const float2* src = base + (size_t)row * P.pitch;
#pragma unroll
        for (int j = 0; j < 10; ++j) {          // lane L of group j: cells 64j + 2L, 64j + 2L + 1 (nc is even)
          const int c = 64 * j + 2 * lane;
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row < w[2] && j < nTW && c < nc) v[j] = ld_stream_f4(reinterpret_cast<const float4*>(src + c));
        }
      };
      float4 cur[10], nxt[10];
      load_row(r0 + warp, cur);
      for (int row = r0 + warp; row < w[2]; row += NWC) {
        load_row(row + NWC, nxt);
        unsigned* Urow = Ubuf0 + (size_t)row * UW;
        unsigned wd[20];
#pragma unroll
        for (int j = 0; j < 10; ++j) {               // visited/total > 0.5 (:29-31); bit i of a plain word = cell i
          const unsigned pq = (2.f * cur[j].x > cur[j].y ? 1u : 0u) | (2.f * cur[j].z > cur[j].w ? 2u : 0u);
          const unsigned lo = __shfl_sync(FULL, pq, lane >> 1), hi = __shfl_sync(FULL, pq, 16 + (lane >> 1));
          wd[2 * j] = __ballot_sync(FULL, (lo >> (lane & 1)) & 1u);
          wd[2 * j + 1] = __ballot_sync(FULL, (hi >> (lane & 1)) & 1u);
        }
        if (lane == 0) {                               // out-of-window cells were loaded as zeros -> zero bits
#pragma unroll
          for (int j = 0; j < 20; j += 4)
            if (j < UW) *reinterpret_cast<uint4*>(Urow + j) = make_uint4(wd[j], wd[j + 1], wd[j + 2], wd[j + 3]);
          for (int t = 20; t < UW; t += 4) *reinterpret_cast<uint4*>(Urow + t) = make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < 10; ++j) cur[j] = nxt[j];
      }
    } else
    for (int row = r0 + warp; row < w[2]; row += NWC) {
      const float2* src = base + (size_t)row * P.pitch;
      unsigned* Urow = Ubuf0 + (size_t)row * UW;
      for (int t0 = 0; t0 < nTW; t0 += 10) {         // 10 16-byte loads (a whole row at c3) in flight per lane
        float4 v[10];                                // lane L of group t: cells 64t + 2L, 64t + 2L + 1
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          const int c = 64 * (t0 + j) + 2 * lane;    // nc is even: both cells are inside the window or neither is
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (t0 + j < nTW && c < nc) v[j] = ld_stream_f4(reinterpret_cast<const float4*>(src + c));
        }
        unsigned wd[20];
#pragma unroll
        for (int j = 0; j < 10; ++j) {               // visited/total > 0.5 (:29-31); bit i of a plain word = cell i
          const unsigned pq = (2.f * v[j].x > v[j].y ? 1u : 0u) | (2.f * v[j].z > v[j].w ? 2u : 0u);
          const unsigned lo = __shfl_sync(FULL, pq, lane >> 1), hi = __shfl_sync(FULL, pq, 16 + (lane >> 1));
          wd[2 * j] = __ballot_sync(FULL, (lo >> (lane & 1)) & 1u);
          wd[2 * j + 1] = __ballot_sync(FULL, (hi >> (lane & 1)) & 1u);
        }
        if (lane == 0) {                               // out-of-window cells were loaded as zeros -> zero bits
#pragma unroll
          for (int j = 0; j < 20; j += 4)
            if (2 * t0 + j < UW) *reinterpret_cast<uint4*>(Urow + 2 * t0 + j) = make_uint4(wd[j], wd[j + 1], wd[j + 2], wd[j + 3]);
        }
      }
      if (lane == 0)
        for (int t = 20 * ((nTW + 9) / 10); t < UW; t += 4) *reinterpret_cast<uint4*>(Urow + t) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (cyc && ctid() == 0) { const long long t = clock64(); cyc[6] += t; cyc[29] += t; }
  }
}

// One particle: wait for its union bitmap, coarse stage, fine stage, results.
__device__ __noinline__ void match_particle(const MatchParams& P, CtaShared& sh, int p, int k) {
  long long* cyc = P.dbgCycles ? P.dbgCycles + (size_t)blockIdx.x * 48 : nullptr;   // [0..15] phases, [16..47] sub-phases
  int status = 0;
  if (cyc && ctid() == 0) { const long long t = clock64(); cyc[6] -= t; if (k == 0) cyc[28] -= t; }
  // this particle's union bitmap is complete: one thread polls the mbarrier, the others park at the named barrier
  // (14 spinning warps would take issue slots from the stream warps they are waiting for)
  if (ctid() == 0) mbar_wait(smem_u32(&sh.ss.uFull[k & 1]), (unsigned)((k >> 1) & 1));
  csync();
  if (cyc && ctid() == 0) { const long long t = clock64(); cyc[6] += t; if (k == 0) cyc[28] += t; }
  if (P.fast) {      // everything but the sparse fine field lives in shared memory
    if (P.fieldOnlyStage != 1) run_stage<true, true>(P, sh, 0, p, k, status);
    if (P.fieldOnlyStage != 0) run_stage<true, false>(P, sh, 1, p, k, status);
  } else {           // large windows: bitmaps / fields / scores in the global slot
    if (P.fieldOnlyStage != 1) run_stage<false, false>(P, sh, 0, p, k, status);
    if (P.fieldOnlyStage != 0) run_stage<false, false>(P, sh, 1, p, k, status);
  }
  status = block_or(status, sh.bs);
  if (P.fieldOnlyStage >= 0) {
    if (ctid() == 0) P.status[p] |= status;
    return;
  }
  if (ctid() == 0) {
    const StageOut c = sh.res[0], f = sh.res[1];
    P.outPose[3 * p] = f.x; P.outPose[3 * p + 1] = f.y; P.outPose[3 * p + 2] = f.th;
    P.outConf[p] = c.conf;
    int* o = P.outIdx + 6 * p;
    o[0] = c.it; o[1] = c.ia; o[2] = c.ib; o[3] = f.it; o[4] = f.ia; o[5] = f.ib;
    P.status[p] |= status;     // OR: bits raised earlier in the step (slam_propose_poses) survive
  }
}

__global__ void __launch_bounds__(NT_ALL, 1) match_kernel(const __grid_constant__ MatchParams P,
                                                           const __grid_constant__ CUtensorMap tmap) {
  __shared__ __align__(16) CtaShared sh;
  if (threadIdx.x == 0) {
    for (int i = 0; i < RING_STAGES; ++i) mbar_init(smem_u32(&sh.ss.ringFull[i]), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&sh.ss.uFull[i]), 32 * NSW); mbar_init(smem_u32(&sh.ss.uEmpty[i]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // warps 0 .. NSW-1: union-window stream (TMA producer + bit packer), runs up to one particle ahead of the compute warps
  if (threadIdx.x < 32 * NSW) {
    stream_role(P, &tmap, sh.ss, sh.uwin);
    return;
  }
  if ((int)blockIdx.x < P.N) cold_start(P);
  int k = 0;
  for (int p = blockIdx.x; p < P.N; p += gridDim.x, ++k) match_particle(P, sh, p, k);
}


// ------------------------------------------------------------------------------------------------ stage-level entries
// ScanMatcher.searchToMatch (:91-151) against a CALLER-PROVIDED dense likelihood field (the reference's public stage
// API passes probSP between frameSearchSpace and searchToMatch).  One CTA walks over particles; not the hot path --
// slam_match_scan keeps the field on chip -- but the same list / pairwise-sum / select code, so convTotal is bit-equal.
struct CorrParams {
  int N, K, stage, sampleMode;
  double fovHalf, maxRange;
  const double *prob;          // [N][probStride], row pitch probPitch
  size_t probStride;
  int probPitch;
  const int* dims;             // [N][2] rows, cols
  const double *ranges, *centre, *origin, *rv, *tw, *uniforms;
  double *vol, *outConf;
  int* outIdx;
  int* status;
  unsigned char* scratch;      // per CTA: nPoses doubles
  size_t slotBytes;
  int Kpad, oLists, oCnt, oDx, oDy, oLeaf, oRed;
};

struct FetchGlobalDense {
  static constexpr bool kWide = false;
  static constexpr bool kKeys8 = false;
  const unsigned* list;
  const double* base;
  unsigned off;
  int pitch, two;
  __device__ __forceinline__ void get(int k, double (&v)[GRP]) const {
    const unsigned sxy = list[k] + off;
    const double* q = base + (size_t)(sxy & 0xffffu) * pitch + (sxy >> 16);
    v[0] = q[0];
    v[1] = two ? q[1] : 0.0;
  }
};

__global__ void __launch_bounds__(256) correlate_kernel(const __grid_constant__ CorrParams C, const __grid_constant__ StageDev S) {
  __shared__ double s_bestV[8];
  __shared__ int s_bestI[8], s_cnt[16], s_K0, s_status, s_chosen;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NWB = blockDim.x >> 5;
  double* dxs = sbuf<double>(C.oDx);
  double* dys = sbuf<double>(C.oDy);
  unsigned* lists = sbuf<unsigned>(C.oLists);
  int* cnts = sbuf<int>(C.oCnt);
  double* leafSum = sbuf<double>(C.oLeaf);
  double* e = reinterpret_cast<double*>(C.scratch + (size_t)blockIdx.x * C.slotBytes);
  const int nOff = S.nOff, nOff2 = nOff * nOff, nPoses = S.nPoses;
  const double ul = S.unitLength;
  for (int p = blockIdx.x; p < C.N; p += gridDim.x) {
    const double cx = C.centre[3 * p], cy = C.centre[3 * p + 1], cth = C.centre[3 * p + 2];
    const double bx = C.origin[2 * p], by = C.origin[2 * p + 1];
    const int Wy = C.dims[2 * p], Wx = C.dims[2 * p + 1];
    int status = 0;
    if (tid == 0) s_status = 0;
    // beam end points (:81-89), order-preserving compaction of beams < maxRange
    {
      const double start = dsub(cth, C.fovHalf), stop = dadd(cth, C.fovHalf);
      const double step = ddiv(dsub(stop, start), (double)(C.K - 1));
      int base = 0;
      for (int k0 = 0; k0 < C.K; k0 += blockDim.x) {
        const int k = k0 + tid;
        bool keep = false;
        double ddx = 0, ddy = 0;
        if (k < C.K) {
          const double rm = C.ranges[k];
          keep = rm < C.maxRange;
          const double ang = (k == C.K - 1) ? stop : dadd(dmul((double)k, step), start);
          double sn, cs;
          sincos(ang, &sn, &cs);
          const double px = dadd(cx, dmul(cs, rm)), py = dadd(cy, dmul(sn, rm));
          ddx = dsub(px, cx); ddy = dsub(py, cy);
        }
        const unsigned bal = __ballot_sync(FULL, keep);
        __syncthreads();
        if (lane == 0) s_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = base, tot = 0;
        for (int w2 = 0; w2 < NWB; ++w2) { if (w2 < warp) before += s_cnt[w2]; tot += s_cnt[w2]; }
        if (keep) { const int pos = before + __popc(bal & ((1u << lane) - 1u)); dxs[pos] = ddx; dys[pos] = ddy; }
        base += tot;
      }
      if (tid == 0) s_K0 = base;
    }
    __syncthreads();
    const int K0 = s_K0;
    for (int tl = warp; tl < S.nTheta; tl += NWB) {
      if (S.E == 8) build_list<8>(dxs, dys, K0, cx, cy, S.cosT[tl], S.sinT[tl], bx, by, ul, S.nHalf, Wx, Wy, lists + tl * C.Kpad, cnts + tl, lane, status);
      else build_list<16>(dxs, dys, K0, cx, cy, S.cosT[tl], S.sinT[tl], bx, by, ul, S.nHalf, Wx, Wy, lists + tl * C.Kpad, cnts + tl, lane, status);
    }
    __syncthreads();
    // scores (:125-132)
    const double* prob = C.prob + (size_t)p * C.probStride;
    double* vol = C.vol + (size_t)p * nPoses;
    const double* tw = C.tw ? C.tw + (size_t)p * nOff2 : nullptr;
    const int nGrp = (nOff + GRP - 1) / GRP, perTheta = nOff * nGrp;
    double best = 0.0;
    int bestIdx = -1, sawNan = 0;
    PruneCtx pc;
    pc.bestS = 0u; pc.h1 = 0; pc.validMask = 0;
    for (int q = tid; q < S.nTheta * perTheta; q += blockDim.x) {
      const int tl = q / perTheta, rem0 = q - tl * perTheta;
      const int a = rem0 / nGrp, b0 = (rem0 - a * nGrp) * GRP;
      FetchGlobalDense f;
      f.list = lists + tl * C.Kpad; f.base = prob; f.pitch = C.probPitch; f.two = b0 + 1 < nOff;
      f.off = (unsigned)(((b0 - S.nHalf) << 16) + (a - S.nHalf));
      double sc[GRP];
      pairwise_g<false>(f, cnts[tl], sc, pc);
      for (int g = 0; g < GRP; ++g) {
        const int b = b0 + g;
        if (b >= nOff) continue;
        const int rem = a * nOff + b, flat = tl * nOff2 + rem;
        double v = sc[g];
        if (C.rv) v = dadd(v, C.rv[rem]);
        if (tw) v = dadd(v, tw[rem]);
        vol[flat] = v;
        if (v != v) sawNan = 1;
        if (first_max_better(v, flat, best, bestIdx)) { best = v; bestIdx = flat; }
      }
    }
    // first maximum in C order
    for (int d = 16; d > 0; d >>= 1) {
      const double ob = __shfl_xor_sync(FULL, best, d);
      const int oi = __shfl_xor_sync(FULL, bestIdx, d);
      if (first_max_better(ob, oi, best, bestIdx)) { best = ob; bestIdx = oi; }
    }
    if (__any_sync(FULL, sawNan) && C.sampleMode) status |= SLAM_ST_NAN_SCORE;      // argmax: numpy returns the first NaN
    if (lane == 0) { s_bestV[warp] = best; s_bestI[warp] = bestIdx; }
    if (status) atomicOr(&s_status, status);
    __syncthreads();
    if (tid == 0) {
      double b = s_bestV[0]; int bi = s_bestI[0];
      for (int w2 = 1; w2 < NWB; ++w2) {
        const double ob = s_bestV[w2]; const int oi = s_bestI[w2];
        if (first_max_better(ob, oi, b, bi)) { b = ob; bi = oi; }
      }
      s_chosen = bi < 0 ? 0 : bi;
    }
    // confidence = np.sum(np.exp(convTotal)): numpy pairwise sum over the flattened volume (:141)
    for (int i = tid; i < nPoses; i += blockDim.x) e[i] = exp(vol[i]);
    __syncthreads();
    for (int g0 = warp * 4; g0 < S.nLeaves; g0 += NWB * 4) {
      const int g = g0 + (lane >> 3), l = lane & 7;
      const bool valid = g < S.nLeaves;
      const int2 lf = valid ? S.leaves[g] : make_int2(0, 0);
      const int off = lf.x, n = lf.y, m = n - (n & 7);
      double r = (n >= 8) ? e[off + l] : 0.0;
      for (int i = 8; i < m; i += 8) r = dadd(r, e[off + i + l]);
      r = dadd(r, __shfl_xor_sync(FULL, r, 1));
      r = dadd(r, __shfl_xor_sync(FULL, r, 2));
      r = dadd(r, __shfl_xor_sync(FULL, r, 4));
      if (n < 8) r = 0.0;
      for (int i = (n < 8 ? 0 : m); i < n; ++i) r = dadd(r, e[off + i]);
      if (valid && l == 0) leafSum[g] = r;
    }
    __syncthreads();
    if (warp == 0) {
      for (int l = 0; l < S.nLevels; ++l) {
        const int o1 = S.levelStart[l + 1];
        for (int o = S.levelStart[l] + lane; o < o1; o += 32) {
          const int2 op = S.ops[o];
          leafSum[S.nLeaves + o] = dadd(leafSum[op.x], leafSum[op.y]);
        }
        __syncwarp();
      }
    }
    __syncthreads();
    const double conf = leafSum[S.nOps ? S.nLeaves + S.nOps - 1 : 0];
    if (C.sampleMode) {   // np.random.choice (:137-139): p = e / e.sum(); cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(u, 'right')
      for (int i = tid; i < nPoses; i += blockDim.x) e[i] = ddiv(e[i], conf);
      __syncthreads();
      if (tid == 0) {
        const double u = C.uniforms[p];
        double c = 0.0;
        for (int i = 0; i < nPoses; ++i) c = dadd(c, e[i]);
        const double last = c;
        c = 0.0;
        int found = nPoses;
        for (int i = 0; i < nPoses; ++i) {
          c = dadd(c, e[i]);
          if (ddiv(c, last) > u) { found = i; break; }
        }
        s_chosen = min(found, nPoses - 1);
      }
      __syncthreads();
    }
    if (tid == 0) {
      const int chosen = s_chosen, it = chosen / nOff2, rem = chosen - it * nOff2;
      C.outIdx[3 * p] = it; C.outIdx[3 * p + 1] = rem / nOff; C.outIdx[3 * p + 2] = rem - (rem / nOff) * nOff;
      C.outConf[p] = conf;
      C.status[p] |= s_status;
    }
    __syncthreads();
  }
}

// scipy.ndimage.gaussian_filter on an arbitrary float64 array (generateProbSearchSpace :41-45): one axis per launch,
// reflect (half-sample symmetric) borders, out = x[c]*w[r]; out += (x[c+j] + x[c-j])*w[j+r], j = -r..-1.
__global__ void blur_axis_kernel(const double* in, double* out, int rows, int cols, int axis, int r, const double* w) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  const int y = (int)(i / cols), x = (int)(i - (size_t)y * cols);
  const int n = axis == 0 ? rows : cols, c = axis == 0 ? y : x;
  auto at = [&](int k) { const int kk = reflect_idx(k, n); return axis == 0 ? in[(size_t)kk * cols + x] : in[(size_t)y * cols + kk]; };
  double v = dmul(at(c), w[r]);
  for (int j = -r; j < 0; ++j) v = dadd(v, dmul(dadd(at(c + j), at(c - j)), w[j + r]));
  out[i] = v;
}
__global__ void min_kernel(const double* in, size_t n, double* out) {       // one block
  __shared__ double s[32];
  double v = INFINITY;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) v = fmin(v, in[i]);
  for (int d = 16; d > 0; d >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, d));
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { for (int k = 1; k < (int)(blockDim.x >> 5); ++k) v = fmin(v, s[k]); out[0] = v; }
}
__global__ void clamp_kernel(double* a, size_t n, const double* probMin) {   // probSP[probSP > 0.5 * probMin] = 0 (:44)
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && a[i] > dmul(0.5, probMin[0])) a[i] = 0.0;
}

// First-pass (axis 0) value for every occupancy pattern of a column's 2r+1 rows; same operation order as scipy:
// out = x[c]*w[r]; out += (x[c+j] + x[c-j])*w[j+r] for j = -r..-1 with x = 0 (occupied) or log(missProb).
__global__ void lut_kernel(double* lut, int r, double C0, const double* T1, const double* T2) {
  const unsigned sr = blockIdx.x * blockDim.x + threadIdx.x;
  if (sr >= (2u << (2 * r))) return;
  double v = ((sr >> r) & 1u) ? 0.0 : C0;
  for (int jj = 0; jj < r; ++jj) {
    const int occ = (int)((sr >> jj) & 1u) + (int)((sr >> (2 * r - jj)) & 1u);
    v = dadd(v, occ == 0 ? T2[jj] : (occ == 1 ? T1[jj] : 0.0));
  }
  lut[sr] = v;
}

// heading prior (ScanMatcher_OGBased.py:105-108)
__global__ void priors_kernel(int N, int nHalf, double coef, const double* phi, const int* hasPhi, double* tw) {
  const int nOff = 2 * nHalf + 1, nOff2 = nOff * nOff;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * nOff2) return;
  const int p = (int)(i / nOff2), rem = (int)(i - (long long)p * nOff2);
  if (!hasPhi[p]) { tw[i] = 0.0; return; }
  const int a = rem / nOff, b = rem - a * nOff;
  const int xv = b - nHalf, yv = a - nHalf;
  double dist = sqrt((double)(xv * xv + yv * yv));
  if (dist == 0.0) dist = 0.0001;
  const double ph = phi[p];
  const double num = dadd(dmul((double)xv, cos(ph)), dmul((double)yv, sin(ph)));
  const double ang = acos(ddiv(num, dist));
  tw[i] = dmul(coef, dmul(ang, ang));
}

}  // namespace slam

// ------------------------------------------------------------------------------------------------ host
using namespace slam;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct slam_matcher {
  MatchParams P;
  std::vector<void*> owned;
  size_t smemBytes;
  int numCtas;
  size_t workspaceBytes;
  int device = -1;           // device the handle was created on (tables live there)
  bool attrSet = false;      // dynamic shared-memory opt-in done for this handle's device
  EncodeTiledFn encode = nullptr;
  alignas(64) CUtensorMap tmap;
  const void* tmapGrid = nullptr;
  int tmapN = 0;
};

// numpy's pairwise recursion over n elements: leaves (offset, length <= 128) and the internal nodes (a, b, height)
static int leaves_rec(int off, int n, std::vector<int2>& leaves, std::vector<int3>& nodes, int& height) {
  if (n <= 128) {
    leaves.push_back(make_int2(off, n));
    height = 0;
    return -(int)leaves.size();                 // leaf k encoded as -(k + 1)
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  int ha = 0, hb = 0;
  const int a = leaves_rec(off, n2, leaves, nodes, ha);
  const int b = leaves_rec(off + n2, n - n2, leaves, nodes, hb);
  height = std::max(ha, hb) + 1;
  nodes.push_back(make_int3(a, b, height));
  return (int)nodes.size() - 1;
}

template <class T>
static int upload(slam_matcher* m, const T* h, size_t n, const T** d) {
  void* p = nullptr;
  SLAM_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  m->owned.push_back(p);
  SLAM_CUDA(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice));
  *d = (const T*)p;
  return 0;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int plan_stage(slam_matcher* m, const slam_geometry* g, const slam_stage_desc& d, double R, int stageId,
                      StageDev& S, size_t& smemNeed, size_t& slotBytes, size_t smemBudget, bool fast, bool uploadTables) {
  if (d.blurRadius < 1 || d.blurRadius > SLAM_MAX_BLUR_RADIUS) return fail(SLAM_E_UNSUPPORTED, "blur radius must be 1..8");
  if (g->K < 2 || g->K > SLAM_MAX_BEAMS) return fail(SLAM_E_UNSUPPORTED, "beams per scan must be 2..512");
  if (d.nHalf < 0 || d.nHalf > 15) return fail(SLAM_E_UNSUPPORTED, "search half-width must be 0..15 cells");
  S.unitLength = d.unitLength;
  S.logMiss = d.logMiss;
  S.r = d.blurRadius;
  const int r = S.r;
  const double c = d.logMiss;
  for (int i = 0; i < 2 * r + 1; ++i) S.w[i] = d.blurW[i];
  // all constants with the same (non-fused) operation order the kernel / scipy use
  volatile double t;
  t = c * S.w[r]; S.C0 = t;
  for (int jj = 0; jj < r; ++jj) {
    t = c * S.w[jj]; S.T1[jj] = t;
    volatile double cc = c + c;
    t = cc * S.w[jj]; S.T2[jj] = t;
  }
  volatile double b1 = S.C0;
  for (int jj = 0; jj < r; ++jj) { volatile double u = b1 + S.T2[jj]; b1 = u; }
  S.B1 = b1;
  volatile double b2 = S.B1 * S.w[r];
  for (int jj = 0; jj < r; ++jj) {
    volatile double pr = S.B1 + S.B1;
    volatile double pm = pr * S.w[jj];
    volatile double u = b2 + pm;
    b2 = u;
  }
  S.B2 = b2;
  S.nHalf = d.nHalf;
  S.nOff = 2 * d.nHalf + 1;
  S.nTheta = d.nTheta;
  S.nPoses = S.nTheta * S.nOff * S.nOff;
  (void)uploadTables;
  if (!S.thetas && (upload(m, d.h_thetas, d.nTheta, &S.thetas) || upload(m, d.h_cos, d.nTheta, &S.cosT) ||
                    upload(m, d.h_sin, d.nTheta, &S.sinT)))
    return 1;
  if (!S.lutV) {
    const size_t n = (size_t)2 << (2 * r);
    void* lut = nullptr;
    SLAM_CUDA(cudaMalloc(&lut, n * sizeof(double)));
    m->owned.push_back(lut);
    const double *dT1 = nullptr, *dT2 = nullptr;
    if (upload(m, S.T1, SLAM_MAX_BLUR_RADIUS, &dT1) || upload(m, S.T2, SLAM_MAX_BLUR_RADIUS, &dT2)) return 1;
    lut_kernel<<<(unsigned)((n + 255) / 256), 256>>>((double*)lut, r, S.C0, dT1, dT2);
    SLAM_CUDA(cudaGetLastError());
    SLAM_CUDA(cudaDeviceSynchronize());
    S.lutV = (const double*)lut;
  }
  std::vector<int2> leaves;
  std::vector<int3> nodes;
  int rootH = 0;
  leaves_rec(0, S.nPoses, leaves, nodes, rootH);
  if (leaves.size() > 30000 || rootH + 1 >= 24) return fail(SLAM_E_UNSUPPORTED, "score volume too large");
  S.nLeaves = (int)leaves.size();
  S.nOps = (int)nodes.size();
  S.nLevels = rootH;
  // internal nodes ordered by height (children first); ids: leaf k -> k, op o -> nLeaves + o
  std::vector<int> order(nodes.size()), newId(nodes.size());
  for (size_t i = 0; i < nodes.size(); ++i) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return nodes[x].z < nodes[y].z; });
  for (size_t i = 0; i < order.size(); ++i) newId[order[i]] = (int)i;
  std::vector<int2> ops(nodes.size());
  for (int l = 0; l < 24; ++l) S.levelStart[l] = (int)nodes.size();
  for (size_t i = 0; i < order.size(); ++i) {
    const int3 nd = nodes[order[i]];
    auto id = [&](int c) { return c < 0 ? -c - 1 : S.nLeaves + newId[c]; };
    ops[i] = make_int2(id(nd.x), id(nd.y));
    S.levelStart[nd.z - 1] = std::min(S.levelStart[nd.z - 1], (int)i);
  }
  for (int l = 22; l >= 0; --l) S.levelStart[l] = std::min(S.levelStart[l], S.levelStart[l + 1]);
  if (!S.leaves && (upload(m, leaves.data(), leaves.size(), &S.leaves) || upload(m, ops.data(), ops.size(), &S.ops)))
    return 1;

  // ---- memory plan
  S.Wmax = (int)(2.0 * R / d.unitLength) + 2;
  S.Wmap = (int)(2.0 * R / g->unit) + 3;
  if (S.Wmax > 32000 || S.Wmap > 32000) return fail(SLAM_E_UNSUPPORTED, "search window too large");
  S.words = (S.Wmax + GRP) / 32 + 2;      // >= 1 always-zero spare word per row (two-word funnel reads)
  S.WT = (S.Wmax + 31) / 32 + 1;
  if (S.WT % 2 == 0) S.WT += 1;          // odd column pitch: conflict-free transposed-bitmap reads
  if ((size_t)S.Wmax * S.words >= 65536) return fail(SLAM_E_UNSUPPORTED, "search window too large (16-bit tile ids)");
  S.Kpad = (g->K + 3) & ~3;
  S.E = g->K <= 256 ? 8 : 16;
  S.needScores = (stageId == 0);
  const size_t bitsBytes = (size_t)S.Wmax * S.words * 4;
  const size_t bitsTBytes = (size_t)S.Wmax * S.WT * 4;
  S.tileCap = std::max(((S.Wmax + 2 * NW - 1) / (2 * NW) + 1) * 2 * S.words, 32 * ((S.Wmax * S.words) / NT + 2));
  const size_t tilesBytes = align_up((size_t)NW * S.tileCap * 2, 16);
  const size_t mapBytes = align_up((size_t)S.Wmap * 2, 16);
  const size_t scoreBytes = S.needScores ? (size_t)S.nPoses * 8 : 0;
  const size_t leafBytes = S.needScores ? (size_t)(S.nLeaves + S.nOps) * 8 : 0;
  const size_t dxyBytes = align_up((size_t)g->K * 8, 16);
  // Field pitch and score-task layout.  A warp's 32 score tasks (same theta, same point k) read GRP adjacent doubles
  // of consecutive offset rows.  For the dense shared-memory field the (pitch, tasks per row) pair is chosen by
  // simulating the bank conflicts of those 64-bit loads (16 double-wide banks, one wavefront per distinct address
  // and bank within a half-warp): padding a row of 7 tasks to 8 lets two rows interleave on even / odd banks.
  int pp = S.Wmax;
  S.nGrpPad = (S.nOff + GRP - 1) / GRP;
  if (stageId == 0 && fast) {
    double bestCost = 1e300;
    const int nGrp0 = (S.nOff + GRP - 1) / GRP;
    for (int pad = nGrp0; pad <= ((nGrp0 + 7) & ~7); ++pad) {
      for (int cand = S.Wmax; cand < S.Wmax + 16; ++cand) {
        const int per = S.nOff * pad;
        long long waves = 0;
        for (int q0 = 0; q0 < per; q0 += 16) {          // half-warps of one theta (theta boundaries ignored)
          int cnt[16][16], nb[16];
          for (int b = 0; b < 16; ++b) nb[b] = 0;
          int mx = 0;
          for (int l = 0; l < 16 && q0 + l < per; ++l) {
            const int q = q0 + l, a = q / pad, b0 = (q - a * pad) * GRP;
            if (b0 >= S.nOff) continue;
            const int addr = a * cand + b0, bank = addr & 15;
            bool dup = false;
            for (int j = 0; j < nb[bank]; ++j) dup |= cnt[bank][j] == addr;
            if (!dup) cnt[bank][nb[bank]++] = addr;
            mx = std::max(mx, nb[bank]);
          }
          waves += mx;
        }
        const double cost = (double)waves + 1e-3 * (cand - S.Wmax);      // ties: the smaller pitch
        if (cost < bestCost) { bestCost = cost; pp = cand; S.nGrpPad = pad; }
      }
    }
  } else if (stageId == 0) {
    while ((pp % 16) != (S.nOff % 16)) ++pp;
  } else {
    pp = (pp + 3) & ~3;
  }
  S.Ppitch = pp;
  const size_t PBytes = (size_t)(S.Wmax + 1) * S.Ppitch * 8;   // one slack row: grouped gathers may overshoot
  // FAST plan: bitmaps, lists, scores (and the dense coarse field) in shared memory; SLOW plan: global slot.
  S.PInSmem = fast ? (stageId == 0) : 0;
  S.scoresInSmem = fast ? 1 : 0;
  S.bitsInSmem = fast ? 1 : 0;
  for (int TB : {S.nTheta, (S.nTheta + 1) / 2, NW, 8, 4}) {
    if (TB > S.nTheta || TB < 1) continue;
    S.TB = TB;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 16); return (int)o; };
    S.oP = S.PInSmem ? take(PBytes) : 0;
    S.oDil = S.bitsInSmem ? take(bitsBytes) : 0;      // activity bitmap lives through the correlation
    S.oDx = take(dxyBytes);
    S.oDy = take(dxyBytes);
    const size_t common = off;
    S.oVw = take((size_t)NW * VW_PER_WARP * 8);       // first-pass values of the blur: aliased by the correlate-phase buffers
    // window / blur-phase buffers
    S.oBits = S.bitsInSmem ? take(bitsBytes) : 0;
    S.oBitsT = S.bitsInSmem ? take(bitsTBytes) : 0;
    S.oRow = take(mapBytes);
    S.oCol = take(mapBytes);
    S.oTiles = S.bitsInSmem ? take(tilesBytes) : 0;    // active-tile ids: one 16-bit id per bitmap word at most, NW segments
    {   // range-path scratch of the scatter: only when small (the coarse stage)
      S.auxPitch = (int)align_up((size_t)std::max(S.Wmax, S.Wmap) + 8, 8);
      const size_t auxBytes = 4 * (size_t)S.auxPitch * 2 + (size_t)S.Wmax * m->P.UW * 4;
      S.auxOK = (auxBytes <= 16384 && off + auxBytes <= smemBudget) ? 1 : 0;
      S.oAux = S.auxOK ? take(auxBytes) : 0;
    }
    const size_t blurEnd = off;
    // correlate-phase buffers alias the blur-phase ones
    off = common;
    S.oLists = take((size_t)TB * S.Kpad * 4);
    S.oCnt = take((size_t)TB * 4);
    S.oScores = (S.needScores && S.scoresInSmem) ? take(scoreBytes) : 0;
    S.oLeaf = take(leafBytes);
    const size_t need = std::max(blurEnd, off);
    if (need <= smemBudget) {
      smemNeed = std::max(smemNeed, need);
      size_t g0 = slotBytes;
      S.gBits = g0; g0 += S.bitsInSmem ? 0 : align_up(bitsBytes, 256);
      S.gBitsT = g0; g0 += S.bitsInSmem ? 0 : align_up(bitsTBytes, 256);
      S.gDil = g0; g0 += S.bitsInSmem ? 0 : align_up(bitsBytes, 256);
      S.gTiles = g0; g0 += S.bitsInSmem ? 0 : align_up(tilesBytes, 256);
      S.gP = g0; g0 += S.PInSmem ? 0 : align_up(PBytes, 256);
      S.gScores = g0; g0 += (S.needScores && !S.scoresInSmem) ? align_up(scoreBytes, 256) : 0;
      slotBytes = g0;
      return 0;
    }
  }
  if (fast) return -1;   // caller retries with the global-slot plan
  return fail(SLAM_E_UNSUPPORTED, "stage does not fit the shared-memory plan");
}

extern "C" int slam_matcher_create(const slam_geometry* g, const slam_matcher_desc* d, slam_matcher** out) {
  if (!g || !d || !out) return fail(SLAM_E_BADARG, "null argument");
  if (g->pitch % 4 || g->pitch < g->G) return fail(SLAM_E_BADARG, "pitch must be a multiple of 4 cells and >= G");
  slam_matcher* m = new slam_matcher();
  MatchParams& P = m->P;
  memset(&P, 0, sizeof(P));
  P.fieldOnlyStage = -1;
  P.G = g->G; P.pitch = g->pitch; P.K = g->K;
  P.unit = g->unit; P.mapX0 = g->mapX0; P.mapX1 = g->mapX1; P.mapY0 = g->mapY0; P.mapY1 = g->mapY1;
  P.fovHalf = g->fovHalf; P.maxRange = g->maxRange; P.R = d->windowRadius;
  P.gridX = g->d_gridX; P.gridY = g->d_gridY;
  int dev = 0, smemMax = 0, sms = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || smemMax <= 0) {
    smemMax = 227 * 1024;   // sm_100a; lets the planner run on a GPU-less build box
    cudaGetLastError();
  }
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    sms = 148;
    cudaGetLastError();
  }
  // union window geometry first: the TMA ring needs RING_STAGES stages of >= 1 window row next to the stage buffers
  const double coarseReach = d->coarse.nHalf * d->coarse.unitLength;
  P.RU = d->windowRadius + coarseReach + 2.0 * g->unit;
  P.UWcells = (((int)(2.0 * P.RU / g->unit) + 8) + 63) / 64 * 64;     // multiple of 64 cells (2 bitmap words)
  P.UW = (2 * (P.UWcells / 64) + 3) / 4 * 4;                       // bitmap words per union row: 16-byte rows
  P.URows = std::min((int)(2.0 * P.RU / g->unit) + 8, g->G);
  {
    const int m64 = P.UWcells / 64;
    const int dsel = groups_per_box(m64);
    P.boxCells = 64 * dsel;                                          // TMA box: <= 256 elements per dimension
    P.nBoxes = m64 / dsel;
  }
  const size_t rowBytes = (size_t)P.UWcells * 8;
  const size_t staticReserve = 4096;                                // static shared memory + driver reserve
  if ((size_t)smemMax < staticReserve + RING_STAGES * rowBytes + 128 + 16384) {
    slam_matcher_destroy(m);
    return fail(SLAM_E_UNSUPPORTED, "union window row too large for the shared-memory TMA ring");
  }
  // the stage planner leaves room for two window rows per ring stage (the stream warps pack rows in pairs)
  const size_t budget = (size_t)smemMax - staticReserve - RING_STAGES * std::min<size_t>(2 * rowBytes, 16384) - 128;
  size_t smemNeed = 0, slot = 0;
  bool fast = true;
  for (int attempt = 0; attempt < 2; ++attempt) {
    smemNeed = 0; slot = 0;
    int rc = 0;
    for (int s = 0; s < 2 && rc == 0; ++s)
      rc = plan_stage(m, g, s == 0 ? d->coarse : d->fine, d->windowRadius, s, P.st[s], smemNeed, slot, budget, fast,
                      attempt == 0);
    if (rc == 0) break;
    if (rc < 0 && fast) { fast = false; continue; }
    slam_matcher_destroy(m);
    return rc;
  }
  P.fast = fast ? 1 : 0;
  // TMA ring after the stage buffers: as many window rows per stage as fit (<= 4), two union bitmaps in the slot
  {
    const size_t avail = (size_t)smemMax - staticReserve - smemNeed - 128;
    int by = (int)(avail / (RING_STAGES * rowBytes));
    by = std::max(1, std::min(by, 4));
    P.ringRows = by;
    P.oRing = (int)align_up(smemNeed, 128);
    smemNeed = (size_t)P.oRing + 128 + (size_t)RING_STAGES * by * rowBytes;
    P.gU = align_up(slot, 256);
    slot = P.gU + 2 * (size_t)P.URows * P.UW * 4;     // double buffered: next particle's bitmap is built early
  }
  m->smemBytes = smemNeed;
  P.slotBytes = align_up(slot, 256);
  m->numCtas = sms;
  m->device = dev;
  m->workspaceBytes = P.slotBytes * (size_t)m->numCtas + 256;
  *out = m;
  return 0;
}

extern "C" void slam_matcher_destroy(slam_matcher* m) {
  if (!m) return;
  for (void* p : m->owned) cudaFree(p);
  delete m;
}

extern "C" size_t slam_matcher_workspace_bytes(const slam_matcher* m) { return m ? m->workspaceBytes : 0; }
extern "C" size_t slam_matcher_workspace_bytes_n(const slam_matcher* m, int32_t N) {
  if (!m || N <= 0) return 0;
  return m->P.slotBytes * (size_t)std::min((int)N, m->numCtas) + 256;      // one slot per CTA that gets a particle
}
extern "C" int slam_matcher_field_side(const slam_matcher* m, int stage) { return m->P.st[stage & 1].Wmax; }
extern "C" int slam_matcher_num_poses(const slam_matcher* m, int stage) { return m->P.st[stage & 1].nPoses; }

// test / profiling hooks (not part of the reference surface)
extern "C" void slam_matcher_set_debug(slam_matcher* m, long long* d_cycles, int flags) {
  m->P.dbgCycles = d_cycles;
  m->P.forceExactCdf = flags & 1;            // bit 0: always walk the CDF sequentially
  m->P.forceGenericScatter = (flags >> 1) & 1;   // bit 1: always take the atomicOr scatter path
  m->P.noPrune = (flags >> 2) & 1;               // bit 2: no branch-and-bound in the fine stage
}
extern "C" int slam_matcher_num_ctas(const slam_matcher* m) { return m->numCtas; }
extern "C" int slam_matcher_plan(const slam_matcher* m, int stage, int* out8) {
  const StageDev& S = m->P.st[stage & 1];
  out8[0] = S.PInSmem; out8[1] = S.scoresInSmem; out8[2] = S.bitsInSmem; out8[3] = S.words;
  out8[4] = S.TB; out8[5] = S.Ppitch; out8[6] = (int)m->smemBytes; out8[7] = (int)(m->P.slotBytes >> 10);
  return 0;
}

static int launch_match(slam_matcher* m, MatchParams& P, int numLattices, void* stream);

extern "C" int slam_match_scan(slam_matcher* m, const float* d_grid, int32_t N, const double* d_ranges,
                               const double* d_estPose, const double* d_rv, const double* d_tw,
                               const double* d_uniforms, double* d_outPose, double* d_outConf, int32_t* d_outIdx,
                               int32_t* d_status, void* d_workspace, size_t workspaceBytes,
                               const slam_match_debug* debug, void* stream) {
  return slam_match_scan_slots(m, d_grid, nullptr, N, N, d_ranges, d_estPose, d_rv, d_tw, d_uniforms, d_outPose, d_outConf,
                               d_outIdx, d_status, d_workspace, workspaceBytes, debug, stream);
}

extern "C" int slam_match_scan_slots(slam_matcher* m, const float* d_grid, const int32_t* d_slots, int32_t numLattices,
                                     int32_t N, const double* d_ranges, const double* d_estPose, const double* d_rv,
                                     const double* d_tw, const double* d_uniforms, double* d_outPose, double* d_outConf,
                                     int32_t* d_outIdx, int32_t* d_status, void* d_workspace, size_t workspaceBytes,
                                     const slam_match_debug* debug, void* stream) {
  if (numLattices < N && !d_slots) return fail(SLAM_E_BADARG, "slam_match_scan_slots: fewer lattices than particles");
  if (!m || !d_grid || !d_ranges || !d_estPose || !d_rv || !d_outPose || !d_outConf || !d_outIdx || !d_status)
    return fail(SLAM_E_BADARG, "slam_match_scan: null argument");
  if (N <= 0) return 0;
  if (!d_workspace || workspaceBytes < slam_matcher_workspace_bytes_n(m, N))
    return fail(SLAM_E_BADARG, "slam_match_scan: workspace too small (slam_matcher_workspace_bytes_n)");
  MatchParams P = m->P;
  P.N = N;
  P.grid = d_grid; P.slots = d_slots; P.ranges = d_ranges; P.estPose = d_estPose; P.rv = d_rv; P.tw = d_tw; P.uniforms = d_uniforms;
  P.outPose = d_outPose; P.outConf = d_outConf; P.outIdx = d_outIdx; P.status = d_status;
  P.scratch = (unsigned char*)align_up((size_t)d_workspace, 256);
  for (int s = 0; s < 2; ++s) {
    P.dbgProb[s] = debug ? debug->d_prob[s] : nullptr;
    P.dbgDims[s] = debug ? debug->d_probDims[s] : nullptr;
    P.dbgVol[s] = debug ? debug->d_vol[s] : nullptr;
    if (P.dbgProb[s] && !P.dbgDims[s]) return fail(SLAM_E_BADARG, "d_prob needs d_probDims");
  }
  return launch_match(m, P, numLattices, stream);
}

static int launch_match(slam_matcher* m, MatchParams& P, int numLattices, void* stream) {
  const float* d_grid = P.grid;
  const int N = P.N;
  int dev = 0;
  SLAM_CUDA(cudaGetDevice(&dev));
  if (dev != m->device) return fail(SLAM_E_BADARG, "slam_match_scan: the current device is not the one the matcher was created on");
  if (!m->attrSet) {
    int optin = 0;
    SLAM_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cudaFuncAttributes fa;
    SLAM_CUDA(cudaFuncGetAttributes(&fa, match_kernel));
    if (m->smemBytes + fa.sharedSizeBytes > (size_t)optin) return fail(SLAM_E_UNSUPPORTED, "slam_match_scan: shared-memory plan exceeds the device limit");
    SLAM_CUDA(cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   optin - (int)fa.sharedSizeBytes));
    m->attrSet = true;
  }
  // TMA descriptor of the lattice batch: [N][G][pitch] cells of 8 bytes, box = boxCells x ringRows x 1
  if (m->tmapGrid != (const void*)d_grid || m->tmapN != numLattices) {
    if (!m->encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      SLAM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      if (!fn || qres != cudaDriverEntryPointSuccess) return fail(SLAM_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available in this driver");
      m->encode = (EncodeTiledFn)fn;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)P.pitch, (cuuint64_t)P.G, (cuuint64_t)numLattices};
    const cuuint64_t strides[2] = {(cuuint64_t)P.pitch * 8, (cuuint64_t)P.pitch * 8 * (cuuint64_t)P.G};
    const cuuint32_t box[3] = {(cuuint32_t)P.boxCells, (cuuint32_t)P.ringRows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult rc = m->encode(&m->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void*)d_grid, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(SLAM_E_UNSUPPORTED, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)rc) + ")");
    m->tmapGrid = d_grid;
    m->tmapN = numLattices;
  }
  const int grid = std::min(N, m->numCtas);
  match_kernel<<<grid, NT_ALL, m->smemBytes, (cudaStream_t)stream>>>(P, m->tmap);
  SLAM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int slam_motion_priors(int32_t N, int32_t nHalf, double coef, const double* d_phi, const int32_t* d_hasPhi,
                                  double* d_tw, void* stream) {
  if (N <= 0) return 0;
  if (!d_phi || !d_hasPhi || !d_tw || nHalf < 0) return fail(SLAM_E_BADARG, "slam_motion_priors: bad argument");
  const long long total = (long long)N * (2 * nHalf + 1) * (2 * nHalf + 1);
  const int bt = 256;
  priors_kernel<<<(unsigned)((total + bt - 1) / bt), bt, 0, (cudaStream_t)stream>>>(N, nHalf, coef, d_phi, d_hasPhi, d_tw);
  SLAM_CUDA(cudaGetLastError());
  return 0;
}

// ---- stage-level entry points: the reference's public frameSearchSpace / searchToMatch /
//      generateProbSearchSpace, for callers that drive the stages themselves
extern "C" int slam_field_build(slam_matcher* m, int32_t stage, const float* d_grid, int32_t N, const double* d_centre,
                                double* d_prob, int32_t* d_probDims, int32_t* d_status, void* d_workspace,
                                size_t workspaceBytes, void* stream) {
  if (!m || !d_grid || !d_centre || !d_prob || !d_probDims || !d_status || (stage != 0 && stage != 1))
    return fail(SLAM_E_BADARG, "slam_field_build: bad argument");
  if (N <= 0) return 0;
  if (!d_workspace || workspaceBytes < slam_matcher_workspace_bytes_n(m, N))
    return fail(SLAM_E_BADARG, "slam_field_build: workspace too small (slam_matcher_workspace_bytes_n)");
  MatchParams P = m->P;
  P.N = N;
  P.fieldOnlyStage = stage;
  P.grid = d_grid; P.slots = nullptr; P.estPose = d_centre; P.status = d_status;
  P.ranges = nullptr; P.rv = nullptr; P.tw = nullptr; P.uniforms = nullptr;
  P.outPose = nullptr; P.outConf = nullptr; P.outIdx = nullptr;
  P.scratch = (unsigned char*)align_up((size_t)d_workspace, 256);
  for (int s = 0; s < 2; ++s) { P.dbgProb[s] = nullptr; P.dbgDims[s] = nullptr; P.dbgVol[s] = nullptr; }
  P.dbgProb[stage] = d_prob;
  P.dbgDims[stage] = d_probDims;
  P.dbgCycles = nullptr;
  return launch_match(m, P, N, stream);
}

extern "C" int slam_correlate(slam_matcher* m, int32_t stage, int32_t N, const double* d_prob, const int32_t* d_probDims,
                              const double* d_ranges, const double* d_centre, const double* d_origin, const double* d_rv,
                              const double* d_tw, const double* d_uniforms, double* d_vol, int32_t* d_outIdx,
                              double* d_outConf, int32_t* d_status, void* d_workspace, size_t workspaceBytes, void* stream) {
  if (!m || !d_prob || !d_probDims || !d_ranges || !d_centre || !d_origin || !d_vol || !d_outIdx || !d_outConf || !d_status ||
      (stage != 0 && stage != 1))
    return fail(SLAM_E_BADARG, "slam_correlate: bad argument");
  if (N <= 0) return 0;
  const StageDev& S = m->P.st[stage];
  CorrParams C;
  memset(&C, 0, sizeof(C));
  C.N = N; C.K = m->P.K; C.stage = stage; C.sampleMode = d_uniforms ? 1 : 0;
  C.fovHalf = m->P.fovHalf; C.maxRange = m->P.maxRange;
  C.prob = d_prob; C.probPitch = S.Wmax; C.probStride = (size_t)S.Wmax * S.Wmax; C.dims = d_probDims;
  C.ranges = d_ranges; C.centre = d_centre; C.origin = d_origin; C.rv = d_rv; C.tw = d_tw; C.uniforms = d_uniforms;
  C.vol = d_vol; C.outIdx = d_outIdx; C.outConf = d_outConf; C.status = d_status;
  C.slotBytes = align_up((size_t)S.nPoses * 8, 256);
  const int ctas = std::min(N, m->numCtas);
  if (!d_workspace || workspaceBytes < C.slotBytes * ctas + 256) return fail(SLAM_E_BADARG, "slam_correlate: workspace too small");
  C.scratch = (unsigned char*)align_up((size_t)d_workspace, 256);
  C.Kpad = S.Kpad;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 16); return (int)o; };
  C.oDx = take((size_t)m->P.K * 8); C.oDy = take((size_t)m->P.K * 8);
  C.oLists = take((size_t)S.nTheta * S.Kpad * 4); C.oCnt = take((size_t)S.nTheta * 4);
  C.oLeaf = take((size_t)(S.nLeaves + S.nOps + 1) * 8);
  int dev = 0, optin = 0;
  SLAM_CUDA(cudaGetDevice(&dev));
  if (dev != m->device) return fail(SLAM_E_BADARG, "slam_correlate: the current device is not the one the matcher was created on");
  SLAM_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (off + 1024 > (size_t)optin) return fail(SLAM_E_UNSUPPORTED, "slam_correlate: lists do not fit shared memory");
  SLAM_CUDA(cudaFuncSetAttribute(correlate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off));
  correlate_kernel<<<ctas, 256, off, (cudaStream_t)stream>>>(C, S);
  SLAM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int slam_blur_clamp(const double* d_in, int32_t rows, int32_t cols, const double* d_taps, int32_t radius,
                               double* d_tmp, double* d_out, void* stream) {
  if (!d_in || !d_taps || !d_tmp || !d_out || rows <= 0 || cols <= 0 || radius < 0)
    return fail(SLAM_E_BADARG, "slam_blur_clamp: bad argument");
  const size_t n = (size_t)rows * cols;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  blur_axis_kernel<<<blocks, 256, 0, st>>>(d_in, d_tmp, rows, cols, 0, radius, d_taps);
  blur_axis_kernel<<<blocks, 256, 0, st>>>(d_tmp, d_out, rows, cols, 1, radius, d_taps);
  min_kernel<<<1, 1024, 0, st>>>(d_out, n, d_tmp);            // d_tmp[0] := probMin
  clamp_kernel<<<blocks, 256, 0, st>>>(d_out, n, d_tmp);
  SLAM_CUDA(cudaGetLastError());
  return 0;
}
