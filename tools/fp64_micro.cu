// Micro-benchmark: FP64 add/mul/fma throughput and latency per SM on this GPU (clock64-based).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP>
__global__ void k(double* out, long long* cyc, double a, double b, int iters) {
  double r[ILP];
  for (int i = 0; i < ILP; ++i) r[i] = a + i + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) r[i] = __dadd_rn(r[i], b);
      if (OP == 1) r[i] = __dmul_rn(r[i], b);
      if (OP == 2) r[i] = __fma_rn(r[i], b, a);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP, int OP>
void run(const char* name, int threads) {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
  int iters = 2000;
  k<ILP, OP><<<148, threads>>>(out, cyc, 1.0000001, 1.0000003, iters);
  k<ILP, OP><<<148, threads>>>(out, cyc, 1.0000001, 1.0000003, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double ops = (double)iters * ILP * threads;
  printf("%-5s ILP=%d threads=%4d : %8.0f cycles, %6.2f thread-ops/clk/SM, %6.1f clk per dependent op\n", name, ILP, threads, c, ops / c, c / iters);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1, 0>("dadd", 32); run<1, 1>("dmul", 32); run<1, 2>("dfma", 32);
  run<8, 0>("dadd", 32); run<8, 0>("dadd", 128); run<8, 0>("dadd", 512); run<8, 0>("dadd", 1024);
  run<8, 1>("dmul", 1024); run<8, 2>("dfma", 1024); run<1, 0>("dadd", 1024); run<2, 0>("dadd", 512);
  return 0;
}
