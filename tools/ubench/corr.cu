// standalone correlation loop: register window (KT+W-1 doubles) x broadcast templates from smem
#include <cstdio>
#include <cuda_runtime.h>
template <int WT, int KT, int VAR>
__device__ __forceinline__ void correlate_reg(const double* __restrict__ t, const double* win, double (&z)[KT]) {
  constexpr int NPAIR = WT / 2;
  double acc[KT][2];
#pragma unroll
  for (int j = 0; j < KT; j++) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
  const double2* t2 = reinterpret_cast<const double2*>(t);
#pragma unroll
  for (int w2 = 0; w2 < NPAIR; w2++) {
    const double2 m = t2[w2];
#pragma unroll
    for (int j = 0; j < KT; j++) {
      acc[j][w2 & 1] = fma(m.x, win[2 * w2 + j], acc[j][w2 & 1]);
      acc[j][w2 & 1] = fma(m.y, win[2 * w2 + 1 + j], acc[j][w2 & 1]);
    }
  }
  const double ml = t[WT - 1];
#pragma unroll
  for (int j = 0; j < KT; j++) z[j] = fma(ml, win[WT - 1 + j], acc[j][NPAIR & 1]) + acc[j][(NPAIR + 1) & 1];
}
template <int KT, int TPB>
__global__ void __launch_bounds__(TPB) k(double* out, const double* in, int G, int reps) {
  extern __shared__ double st[];
  for (int i = threadIdx.x; i < G * 56; i += TPB) st[i] = in[i % 1024] * 1e-2;
  __syncthreads();
  double win[55 + KT - 1];
#pragma unroll
  for (int e = 0; e < 55 + KT - 1; e++) win[e] = in[(threadIdx.x * KT + e) % 1024];
  double s[KT];
#pragma unroll
  for (int j = 0; j < KT; j++) s[j] = 0;
  for (int r = 0; r < reps; r++)
    for (int g = 0; g < G; g++) {
      double z[KT];
      correlate_reg<55, KT, 0>(st + g * 56, win, z);
#pragma unroll
      for (int j = 0; j < KT; j++) s[j] += z[j];
    }
  double tot = 0;
#pragma unroll
  for (int j = 0; j < KT; j++) tot += s[j];
  out[blockIdx.x * TPB + threadIdx.x] = tot;
}
template <int KT, int TPB> void run(double* out, double* in, int blocks_per_sm) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int G = 92, reps = 40;
  size_t smem = G * 56 * 8;
  cudaFuncSetAttribute(k<KT, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int blocks = 148 * blocks_per_sm;
  k<KT, TPB><<<blocks, TPB, smem>>>(out, in, G, 2);
  cudaEventRecord(e0); k<KT, TPB><<<blocks, TPB, smem>>>(out, in, G, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = 2.0 * 55 * KT * G * reps * (double)TPB * blocks;
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<KT, TPB>);
  printf("KT %d TPB %d blocks/SM %d regs %d: %.3f ms  %.2f TFLOP/s (%.0f%% of 37.1)\n", KT, TPB, blocks_per_sm, fa.numRegs, ms, fl / ms / 1e9, fl / ms / 1e9 / 37.1 * 100);
}
int main() {
  double *out, *in; cudaMalloc(&out, 8 * 512 * 148 * 8); cudaMalloc(&in, 8 * 1024);
  double h[1024]; for (int i = 0; i < 1024; i++) h[i] = 1.0 + i * 1e-3; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<4, 128>(out, in, 1); run<4, 128>(out, in, 2); run<4, 128>(out, in, 3);
  run<2, 128>(out, in, 2); run<2, 128>(out, in, 3); run<2, 128>(out, in, 4);
  run<8, 128>(out, in, 1); run<8, 128>(out, in, 2);
  return 0;
}
