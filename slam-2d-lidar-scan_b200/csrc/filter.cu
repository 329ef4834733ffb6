// Particle-filter bookkeeping kernels (Algorithm/FastSlam.py:30-62, :77-135), sm_100a.
// All of this is O(N) scalar float64 work whose summation ORDER is part of the contract (sequential loops in
// the reference), so the order-sensitive parts run in a single thread; N <= 16384 keeps that under ~100 us.
#include "common.cuh"

namespace slam {

// updateEstimatedPose (FastSlam.py:77-106), per-particle part
__global__ void propose_kernel(int N, const double* prevMatched, double rawTheta, double prevRawTheta, int mode,
                               double rawTurn, const double* prevHeading, const int* hasHeading, double* estPose,
                               double* phi, int* hasPhi, int* status) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  estPose[3 * p] = prevMatched[3 * p];
  estPose[3 * p + 1] = prevMatched[3 * p + 1];
  estPose[3 * p + 2] = dsub(dadd(prevMatched[3 * p + 2], rawTheta), prevRawTheta);   // :78, left to right
  int has = 0;
  double ph = 0.0;
  if (mode == 1) {
    if (hasHeading[p]) { has = 1; ph = dadd(prevHeading[p], rawTurn); }               // :94-95
    else atomicOr(&status[p], SLAM_ST_HEADING_MISSING);
  }
  phi[p] = ph;
  hasPhi[p] = has;
}

// getMovingTheta (:108-120) from the last trajectory point, prevMatched := matched, weight *= confidence (:135)
__global__ void finish_kernel(int N, const double* matched, const double* conf, double* prevMatched, double* prevHeading,
                              int* hasHeading, double* weights) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const double x = matched[3 * p], y = matched[3 * p + 1];
  const double mxv = dsub(x, prevMatched[3 * p]), myv = dsub(y, prevMatched[3 * p + 1]);
  const double move = sqrt(dadd(dmul(mxv, mxv), dmul(myv, myv)));
  if (move != 0.0) {
    const double a = acos(ddiv(mxv, move));
    prevHeading[p] = myv > 0.0 ? a : -a;
    hasHeading[p] = 1;
  } else {
    hasHeading[p] = 0;
  }
  prevMatched[3 * p] = x; prevMatched[3 * p + 1] = y; prevMatched[3 * p + 2] = matched[3 * p + 2];
  weights[p] = dmul(weights[p], conf[p]);
}

// normalizeWeights (:43-48) + weightUnbalanced trigger (:32-37).  The two accumulations are sequential float64 by
// contract (one thread adds, in index order); the independent per-element work (loads, divisions, squares) is done
// by the whole block through a shared-memory tile so the adding thread never waits on memory or on a divide.
constexpr int NORM_T = 1024;
// `src` may differ from `w` (out-of-place: the raw weights stay readable); with `status` the kernel also ORs the nStatus
// status words into the low word of out[2] -- the whole end-of-step trigger in one launch.
__global__ void __launch_bounds__(NORM_T) normalize_kernel(int N, const double* src, double* w, double* out,
                                                           const int* status, int nStatus) {
  __shared__ double tile[NORM_T];
  __shared__ double s_sum;
  __shared__ int s_or;
  const int tid = threadIdx.x;
  if (status) {
    if (tid == 0) s_or = 0;
    __syncthreads();
    int v = 0;
    for (int i = tid; i < nStatus; i += NORM_T) v |= status[i];
    v = __reduce_or_sync(0xffffffffu, v);
    if ((tid & 31) == 0 && v) atomicOr(&s_or, v);
  }
  double s = 0.0;
  for (int i0 = 0; i0 < N; i0 += NORM_T) {
    if (i0 + tid < N) tile[tid] = src[i0 + tid];
    __syncthreads();
    if (tid == 0) {
      const int n = min(NORM_T, N - i0);
      for (int i = 0; i < n; ++i) s = dadd(s, tile[i]);
    }
    __syncthreads();
  }
  if (tid == 0) s_sum = s;
  __syncthreads();
  s = s_sum;
  const double n = (double)N;
  const double invN = ddiv(1.0, n);
  double var = 0.0;
  for (int i0 = 0; i0 < N; i0 += NORM_T) {
    if (i0 + tid < N) {
      const double wi = ddiv(src[i0 + tid], s);
      w[i0 + tid] = wi;
      const double d = dsub(wi, invN);
      tile[tid] = dmul(d, d);
    }
    __syncthreads();
    if (tid == 0) {
      const int m = min(NORM_T, N - i0);
      for (int i = 0; i < m; ++i) var = dadd(var, tile[i]);
    }
    __syncthreads();
  }
  if (tid == 0) {
    // ((N-1)/N)**2 + (N - 1.000000000000001) * (1/N)**2
    const double a = ddiv(n - 1.0, n);
    const double thr = dadd(dmul(a, a), dmul(dsub(n, 1.000000000000001), dmul(invN, invN)));
    out[0] = var;
    out[1] = var > thr ? 1.0 : 0.0;
    if (status) reinterpret_cast<int*>(out + 2)[0] = s_or;      // published by the barriers of the loops above
  }
}

// bitwise OR of the per-particle status words (one block)
__global__ void status_or_kernel(int N, const int* status, int* out) {
  __shared__ int s_or;
  if (threadIdx.x == 0) s_or = 0;
  __syncthreads();
  int v = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) v |= status[i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0 && v) atomicOr(&s_or, v);
  __syncthreads();
  if (threadIdx.x == 0) out[0] = s_or;
}

// legacy RandomState.choice: sequential cumsum, normalise by the last element, searchsorted side='right'
__global__ void cdf_kernel(int N, const double* w, double* cdf) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double c = 0.0;
  for (int i = 0; i < N; ++i) { c = dadd(c, w[i]); cdf[i] = c; }
  const double last = c;
  for (int i = 0; i < N; ++i) cdf[i] = ddiv(cdf[i], last);
}

__global__ void search_kernel(int N, const double* cdf, const double* u, int* idx) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const double v = u[p];
  int lo = 0, hi = N;             // first i with cdf[i] > v
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] <= v) lo = mid + 1; else hi = mid;
  }
  idx[p] = min(lo, N - 1);
}

// bulk copy of whole lattices: dst[i] = src[idx[i]] (the deepcopy of FastSlam.py:61); 16-byte vectors
__global__ void gather_grid_kernel(const float4* src, float4* dst, const int* idx, size_t n4PerParticle) {
  const int i = blockIdx.y;
  const float4* s = src + (size_t)idx[i] * n4PerParticle;
  float4* d = dst + (size_t)i * n4PerParticle;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n4PerParticle; k += (size_t)gridDim.x * blockDim.x)
    d[k] = s[k];
}

// in-place copies of a copy-elided resample: lattice dst[c] := lattice src[c] (disjoint sets), 16-byte vectors
__global__ void copy_lattice_kernel(float4* grid, const int* src, const int* dst, size_t n4PerParticle) {
  const int c = blockIdx.y;
  const float4* s = grid + (size_t)src[c] * n4PerParticle;
  float4* d = grid + (size_t)dst[c] * n4PerParticle;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n4PerParticle; k += (size_t)gridDim.x * blockDim.x)
    d[k] = __ldcs(s + k);
}

__global__ void gather_state_kernel(int N, const int* idx, const double* src, double* dst, int cols, double* weights) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  for (int c = 0; c < cols; ++c) dst[(size_t)p * cols + c] = src[(size_t)idx[p] * cols + c];
  weights[p] = ddiv(1.0, (double)N);
}

}  // namespace slam

using namespace slam;

extern "C" int slam_propose_poses(int32_t N, const double* d_prevMatched, double rawTheta, double prevRawTheta,
                                  int32_t mode, double rawTurn, const double* d_prevHeading,
                                  const int32_t* d_hasHeading, double* d_estPose, double* d_phi, int32_t* d_hasPhi,
                                  int32_t* d_status, void* stream) {
  if (N <= 0) return 0;
  if (!d_prevMatched || !d_prevHeading || !d_hasHeading || !d_estPose || !d_phi || !d_hasPhi || !d_status)
    return fail(SLAM_E_BADARG, "slam_propose_poses: null argument");
  propose_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(N, d_prevMatched, rawTheta, prevRawTheta, mode,
                                                                    rawTurn, d_prevHeading, d_hasHeading, d_estPose,
                                                                    d_phi, d_hasPhi, d_status);
  SLAM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int slam_finish_step(int32_t N, const double* d_matched, const double* d_conf, double* d_prevMatched,
                                double* d_prevHeading, int32_t* d_hasHeading, double* d_weights, void* stream) {
  if (N <= 0) return 0;
  if (!d_matched || !d_conf || !d_prevMatched || !d_prevHeading || !d_hasHeading || !d_weights)
    return fail(SLAM_E_BADARG, "slam_finish_step: null argument");
  finish_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(N, d_matched, d_conf, d_prevMatched, d_prevHeading,
                                                                   d_hasHeading, d_weights);
  SLAM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int slam_normalize_weights(int32_t N, double* d_weights, double* d_out, void* stream) {
  if (N <= 0 || !d_weights || !d_out) return fail(SLAM_E_BADARG, "slam_normalize_weights: bad argument");
  normalize_kernel<<<1, NORM_T, 0, (cudaStream_t)stream>>>(N, d_weights, d_weights, d_out, nullptr, 0);
  SLAM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int slam_step_trigger(int32_t N, const double* d_weightsIn, double* d_weightsOut, const int32_t* d_status,
                                 int32_t nStatus, double* d_out, void* stream) {
  if (N <= 0 || !d_weightsIn || !d_weightsOut || !d_out || (nStatus > 0 && !d_status))
    return fail(SLAM_E_BADARG, "slam_step_trigger: bad argument");
  normalize_kernel<<<1, NORM_T, 0, (cudaStream_t)stream>>>(N, d_weightsIn, d_weightsOut, d_out,
                                                           nStatus > 0 ? d_status : nullptr, nStatus);
  SLAM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int slam_status_reduce(int32_t N, const int32_t* d_status, int32_t* d_out, void* stream) {
  if (N <= 0 || !d_status || !d_out) return fail(SLAM_E_BADARG, "slam_status_reduce: bad argument");
  status_or_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(N, d_status, d_out);
  SLAM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int slam_resample_indices(int32_t N, const double* d_weights, const double* d_uniforms, double* d_cdfScratch,
                                     int32_t* d_idx, void* stream) {
  if (N <= 0 || !d_weights || !d_uniforms || !d_cdfScratch || !d_idx)
    return fail(SLAM_E_BADARG, "slam_resample_indices: bad argument");
  cdf_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(N, d_weights, d_cdfScratch);
  search_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(N, d_cdfScratch, d_uniforms, d_idx);
  SLAM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int slam_gather_particles(const slam_geometry* g, int32_t N, const int32_t* d_idx, const float* d_gridSrc,
                                     float* d_gridDst, const double* d_stateSrc, double* d_stateDst, int32_t stateCols,
                                     double* d_weights, void* stream) {
  if (!g || N <= 0 || !d_idx || !d_gridSrc || !d_gridDst || !d_weights || d_gridSrc == d_gridDst)
    return fail(SLAM_E_BADARG, "slam_gather_particles: bad argument");
  const size_t n4 = (size_t)g->G * g->pitch / 2;
  dim3 grid(64, N);
  gather_grid_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)d_gridSrc, (float4*)d_gridDst, d_idx, n4);
  SLAM_CUDA(cudaGetLastError());
  if (stateCols > 0 && d_stateSrc && d_stateDst) {
    gather_state_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(N, d_idx, d_stateSrc, d_stateDst, stateCols,
                                                                           d_weights);
  } else {
    gather_state_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(N, d_idx, nullptr, nullptr, 0, d_weights);
  }
  SLAM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int slam_copy_lattices(const slam_geometry* g, float* d_grid, int32_t nCopies, const int32_t* d_src,
                                  const int32_t* d_dst, void* stream) {
  if (nCopies <= 0) return 0;
  if (!g || !d_grid || !d_src || !d_dst) return fail(SLAM_E_BADARG, "slam_copy_lattices: bad argument");
  const size_t n4 = (size_t)g->G * g->pitch / 2;
  for (int32_t c0 = 0; c0 < nCopies; c0 += 65535) {          // gridDim.y limit
    const int32_t nc = nCopies - c0 < 65535 ? nCopies - c0 : 65535;
    dim3 grid(64, nc);
    copy_lattice_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((float4*)d_grid, d_src + c0, d_dst + c0, n4);
    SLAM_CUDA(cudaGetLastError());
  }
  return 0;
}
