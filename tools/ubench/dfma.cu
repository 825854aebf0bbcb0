// DFMA throughput vs number of fresh register operands (sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, const double* in, int iters) {
  double a[8], b[8], c[8];
  for (int i = 0; i < 8; i++) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i] * 1e-3; c[i] = in[threadIdx.x + 16 + i] * 1e-3; }
  double x = in[0] * 1e-3, y = in[1] * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (MODE == 1) a[i] = fma(a[i], x, y);           // 1 fresh operand
        if (MODE == 2) a[i] = fma(b[i], x, a[i]);        // 2 fresh
        if (MODE == 3) a[i] = fma(b[i], c[(i + r) & 7], a[i]);  // 3 fresh
        if (MODE == 4) a[i] = fma(b[(i + r) & 7], c[r], a[i]);  // corr-like: c[r] shared by 8 consecutive
        if (MODE == 5) a[i & 3] = fma(b[(i + r) & 7], c[(2 * r + (i >> 2)) & 7], a[i & 3]);  // shared by 4, 4 accumulators
      }
    }
  }
  double s = 0; for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(double* out, double* in, int blocks) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096;
  k<MODE><<<blocks, 256>>>(out, in, 64);
  cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = 2.0 * 64 * iters * 256.0 * blocks;
  printf("mode %d: %.3f ms  %.2f TFLOP/s\n", MODE, ms, fl / ms / 1e9);
}
int main() {
  double *out, *in; cudaMalloc(&out, 8 * 256 * 148 * 8); cudaMalloc(&in, 8 * 1024);
  double h[1024]; for (int i = 0; i < 1024; i++) h[i] = 1.0 + i * 1e-3; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int occ = 2; occ <= 8; occ *= 2) { printf("blocks/SM %d\n", occ); run<1>(out, in, 148 * occ); run<2>(out, in, 148 * occ); run<3>(out, in, 148 * occ); run<4>(out, in, 148 * occ); run<5>(out, in, 148 * occ); }
  return 0;
}
