// Microbenchmark: issue-slot model of the float64 hot loops on B200.  A DFMA occupies the 16-lane float64 pipe of an SM
// sub-partition for 2 clocks; do integer / shuffle / shared-memory instructions of the same warps issue "in its shadow",
// or does every instruction cost its own slot (time = 2 x float64 + 1 x other)?  Same question for the float64 tensor-core
// MMA (mma.sync.m8n8k4.f64 = 256 FMAs, 16 pipe clocks).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix2 issue_mix2.cu && ./issue_mix2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// per unrolled step: NF DFMA (independent chains), NM DMMA, NI integer ops (independent chains), NS shuffles, NL LDS
template <int NF, int NM, int NI, int NS, int NL>
__global__ void __launch_bounds__(384) k(double* out, int iters, double a, double b, int m) {
  __shared__ double sh[256];
  if (threadIdx.x < 256) sh[threadIdx.x] = threadIdx.x;
  __syncthreads();
  double v[8], c[8], s = 0.0;
  int x[8];
  float fs[4];
#pragma unroll
  for (int q = 0; q < 8; q++) { v[q] = threadIdx.x + q; c[q] = threadIdx.x - q; x[q] = threadIdx.x + q; }
#pragma unroll
  for (int q = 0; q < 4; q++) fs[q] = threadIdx.x + q;
  const double fa = a + 1e-12 * threadIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int q = 0; q < NF; q++) v[q & 7] = fma(v[q & 7], a, b);
#pragma unroll
      for (int q = 0; q < NM; q++) dmma(c[2 * (q & 3)], c[2 * (q & 3) + 1], fa, b);
#pragma unroll
      for (int q = 0; q < NI; q++) x[q & 7] = (x[q & 7] ^ m) + u;            // LOP3 + IADD -> counted as 2 ops? (one LOP3.LUT + one IADD)
#pragma unroll
      for (int q = 0; q < NS; q++) fs[q & 3] = __shfl_xor_sync(0xffffffffu, fs[q & 3], 1 + (q & 3));
#pragma unroll
      for (int q = 0; q < NL; q++) s += sh[(x[q & 7] + q) & 255];
    }
  }
  double r = s;
#pragma unroll
  for (int q = 0; q < 8; q++) r += v[q] + c[q] + x[q];
#pragma unroll
  for (int q = 0; q < 4; q++) r += fs[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int NF, int NM, int NI, int NS, int NL>
void run(const char* name, int sms, double mhz) {
  double* buf; cudaMalloc(&buf, sizeof(double) * sms * 384);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096; float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0); k<NF, NM, NI, NS, NL><<<sms, 384>>>(buf, iters, 0.999999, 1e-9, 3); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
  }
  const double clk = best * 1e-3 * mhz * 1e6 / (iters * 4.0) / 3.0;   // clocks per unrolled step per warp (3 warps per sub-partition)
  printf("%-52s %8.3f ms  %6.2f clk per step per warp-slot | model 2*DFMA + 16*DMMA = %d, + others = %d\n", name, best, clk,
         2 * NF + 16 * NM, 2 * NF + 16 * NM + 2 * NI + NS + 2 * NL);
  cudaFree(buf);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount; const double mhz = p.clockRate / 1000.0;
  printf("clock %.0f MHz; 1 CTA x 384 threads per SM (3 warps per sub-partition)\n", mhz);
  run<8, 0, 0, 0, 0>("8 DFMA", sms, mhz);
  run<8, 0, 4, 0, 0>("8 DFMA + 4 x (LOP3 + IADD)", sms, mhz);
  run<8, 0, 8, 0, 0>("8 DFMA + 8 x (LOP3 + IADD)", sms, mhz);
  run<8, 0, 0, 4, 0>("8 DFMA + 4 SHFL", sms, mhz);
  run<8, 0, 0, 8, 0>("8 DFMA + 8 SHFL", sms, mhz);
  run<8, 0, 0, 0, 4>("8 DFMA + 4 x (LDS + DADD + LOP/IADD)", sms, mhz);
  run<0, 0, 8, 0, 0>("8 x (LOP3 + IADD) only", sms, mhz);
  run<0, 0, 0, 8, 0>("8 SHFL only", sms, mhz);
  run<0, 1, 0, 0, 0>("1 DMMA", sms, mhz);
  run<0, 2, 0, 0, 0>("2 DMMA", sms, mhz);
  run<0, 2, 4, 0, 0>("2 DMMA + 4 x (LOP3 + IADD)", sms, mhz);
  run<0, 2, 8, 0, 0>("2 DMMA + 8 x (LOP3 + IADD)", sms, mhz);
  run<0, 2, 16, 0, 0>("2 DMMA + 16 x (LOP3 + IADD)", sms, mhz);
  run<0, 2, 0, 8, 0>("2 DMMA + 8 SHFL", sms, mhz);
  run<8, 1, 0, 0, 0>("8 DFMA + 1 DMMA", sms, mhz);
  run<8, 1, 8, 0, 0>("8 DFMA + 1 DMMA + 8 x (LOP3 + IADD)", sms, mhz);
  return 0;
}
