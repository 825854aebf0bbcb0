// DFMA latency / per-warp ILP scaling on sm_100a: 1 warp per SMSP (block of 128 threads, 1 block/SM)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void __launch_bounds__(128) k(double* out, const double* in, int iters) {
  double a[ILP];
  for (int i = 0; i < ILP; i++) a[i] = in[threadIdx.x + i];
  double x = in[0] * 1e-3, y = in[1] * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++)
#pragma unroll
      for (int i = 0; i < ILP; i++) a[i] = fma(a[i], x, y);
  }
  double s = 0; for (int i = 0; i < ILP; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> void run(double* out, double* in, int warps_per_smsp) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2048;
  const int blocks = 148 * warps_per_smsp;
  k<ILP><<<blocks, 128>>>(out, in, 16);
  cudaEventRecord(e0); k<ILP><<<blocks, 128>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double n_dfma_per_warp = 16.0 * ILP * iters;
  double cyc = ms * 1e-3 * 1.965e9;
  printf("warps/SMSP %d ILP %2d: %.3f ms  cycles per DFMA per warp %.2f  (pipe util %.0f%%)\n", warps_per_smsp, ILP, ms, cyc / n_dfma_per_warp,
         100.0 * 2.0 * n_dfma_per_warp * warps_per_smsp / cyc);
}
int main() {
  double *out, *in; cudaMalloc(&out, 8 * 128 * 148 * 8 * 8); cudaMalloc(&in, 8 * 1024);
  double h[1024]; for (int i = 0; i < 1024; i++) h[i] = 1.0 + i * 1e-3; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int w = 1; w <= 4; w *= 2) { run<1>(out, in, w); run<2>(out, in, w); run<4>(out, in, w); run<8>(out, in, w); run<16>(out, in, w); }
  return 0;
}
