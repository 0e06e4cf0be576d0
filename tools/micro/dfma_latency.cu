// Microbenchmark: float64 FMA dependent-issue latency and the parallelism needed to saturate the pipe.
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, int iters, double a, double b) {
  double v[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) v[c] = threadIdx.x + c;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++)
#pragma unroll
      for (int c = 0; c < CH; c++) v[c] = fma(v[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += v[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
void run(int sms, int threads, int blocks_per_sm, double mhz) {
  double* buf; cudaMalloc(&buf, sizeof(double) * sms * blocks_per_sm * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096; float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0); k<CH><<<sms * blocks_per_sm, threads>>>(buf, iters, 0.999999, 1e-9); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
  }
  double per_warp_instr = 16.0 * CH * iters;
  double clk = best * 1e-3 * mhz * 1e6;
  int warps_per_smsp = threads * blocks_per_sm / 128;
  printf("chains %d, warps/SMSP %d: %.3f ms, %.2f clk per DFMA per warp, pipe use %.0f%%\n", CH, warps_per_smsp, best,
         clk / per_warp_instr, 100.0 * per_warp_instr * warps_per_smsp * 2.0 / clk);
  cudaFree(buf);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; double mhz = p.clockRate / 1000.0;
  printf("clock %.0f MHz\n", mhz);
  run<1>(sms, 128, 1, mhz); run<2>(sms, 128, 1, mhz); run<4>(sms, 128, 1, mhz); run<8>(sms, 128, 1, mhz);
  run<1>(sms, 512, 1, mhz); run<2>(sms, 512, 1, mhz); run<4>(sms, 512, 1, mhz);
  run<1>(sms, 512, 2, mhz); run<2>(sms, 512, 2, mhz);
  return 0;
}
