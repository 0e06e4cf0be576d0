// Microbenchmark: does a float64 FMA stream leave issue slots for integer / shared-memory instructions?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu && ./issue_mix
#include <cstdio>
#include <cuda_runtime.h>
template <int NI, int NL>
__global__ void __launch_bounds__(512) k(double* out, int iters, double a, double b, int m) {
  __shared__ double sh[64];
  if (threadIdx.x < 64) sh[threadIdx.x] = threadIdx.x;
  __syncthreads();
  double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
  int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
  double s = 0;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
      v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
#pragma unroll
      for (int q = 0; q < NI; q++) {   // NI * 2 integer ops per 8 DFMA... scaled below
        i0 = (i0 ^ m) + u; i1 = (i1 ^ m) + u; i2 = (i2 ^ m) + u; i3 = (i3 ^ m) + u;
      }
#pragma unroll
      for (int q = 0; q < NL; q++) s += sh[(i + u + q) & 63];   // broadcast LDS (+1 DADD each)
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7)) + (i0 + i1 + i2 + i3) + s;
}
template <int NI, int NL>
void run(const char* name, int sms) {
  double* buf; cudaMalloc(&buf, sizeof(double) * sms * 4 * 512);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2048; float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0); k<NI, NL><<<sms * 4, 512>>>(buf, iters, 0.999999, 1e-9, 3); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
  }
  double dfma = 64.0 * iters * sms * 4 * 512;
  printf("%-34s %.3f ms  %.2f T DFMA/s  (int ops per DFMA %.2f, LDS per DFMA %.3f)\n", name, best, dfma / best / 1e9,
         NI * 8.0 / 8.0, NL / 8.0);
  cudaFree(buf);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  run<0, 0>("dfma only", sms);
  run<1, 0>("dfma + 1.0 int/dfma", sms);
  run<2, 0>("dfma + 2.0 int/dfma", sms);
  run<0, 1>("dfma + 1/8 lds(+dadd)/dfma", sms);
  run<0, 4>("dfma + 1/2 lds(+dadd)/dfma", sms);
  run<1, 2>("dfma + 1 int + 1/4 lds", sms);
  return 0;
}
