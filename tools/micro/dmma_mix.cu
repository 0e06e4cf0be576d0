// Microbenchmark: float64 tensor-core MMA (mma.sync.m8n8k4.f64, "DMMA") on B200 -- its throughput alone, and whether it
// overlaps with a vector DFMA stream (separate pipe) or competes with it (same pipe).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_mix dmma_mix.cu && ./dmma_mix
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// NM independent DMMA accumulator pairs and NF independent DFMA chains per thread, per unrolled iteration.
template <int NM, int NF>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b) {
  double c[2 * (NM > 0 ? NM : 1)], v[NF > 0 ? NF : 1];
#pragma unroll
  for (int q = 0; q < (NM > 0 ? NM : 1); q++) { c[2 * q] = threadIdx.x + q; c[2 * q + 1] = threadIdx.x - q; }
#pragma unroll
  for (int q = 0; q < (NF > 0 ? NF : 1); q++) v[q] = threadIdx.x + 0.5 * q;
  double fa = a + 1e-12 * threadIdx.x, fb = b;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int q = 0; q < NM; q++) dmma(c[2 * q], c[2 * q + 1], fa, fb);
#pragma unroll
      for (int q = 0; q < NF; q++) v[q] = fma(v[q], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int q = 0; q < (NM > 0 ? NM : 1); q++) s += c[2 * q] + c[2 * q + 1];
#pragma unroll
  for (int q = 0; q < (NF > 0 ? NF : 1); q++) s += v[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NM, int NF>
void run(const char* name, int sms, int ctas_per_sm) {
  double* buf; cudaMalloc(&buf, sizeof(double) * sms * ctas_per_sm * 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096; float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0); k<NM, NF><<<sms * ctas_per_sm, 256>>>(buf, iters, 0.999999, 1e-9); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
  }
  double warps = (double)sms * ctas_per_sm * 8;
  double mma_fma = 256.0 * NM * 4 * iters * warps;          // FMAs executed by DMMA (8x8x4 per warp instruction)
  double vec_fma = 32.0 * NF * 4 * iters * warps;           // FMAs executed by DFMA
  printf("%-40s ctas/sm %d  %.3f ms  DMMA %.2f TFLOP/s  DFMA %.2f TFLOP/s  sum %.2f\n", name, ctas_per_sm, best,
         2 * mma_fma / best / 1e9, 2 * vec_fma / best / 1e9, 2 * (mma_fma + vec_fma) / best / 1e9);
  cudaFree(buf);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  for (int c = 1; c <= 4; c *= 2) {
    run<0, 8>("dfma only (8 chains)", sms, c);
    run<4, 0>("dmma only (4 accumulators)", sms, c);
    run<8, 0>("dmma only (8 accumulators)", sms, c);
    run<1, 8>("1 dmma : 8 dfma  (256 : 256 FMA)", sms, c);
    run<2, 8>("2 dmma : 8 dfma  (512 : 256 FMA)", sms, c);
    run<1, 16>("1 dmma : 16 dfma (256 : 512 FMA)", sms, c);
    run<4, 8>("4 dmma : 8 dfma  (1024 : 256 FMA)", sms, c);
  }
  return 0;
}
