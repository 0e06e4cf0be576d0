// Microbenchmark: DFMA throughput vs number of distinct REGISTER operands (constant-bank operands are free).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(double* out, const double* in, int iters) {
  double v[8], w[8], z[8];
#pragma unroll
  for (int c = 0; c < 8; c++) { v[c] = in[threadIdx.x + c]; w[c] = in[threadIdx.x + 8 + c]; z[c] = in[threadIdx.x + 16 + c]; }
  const double a = 0.999999, b = 1e-9;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int c = 0; c < 8; c++) {
        if (MODE == 0) v[c] = fma(v[c], a, b);                 // 1 register operand (+dst)
        if (MODE == 1) v[c] = fma(v[c], w[c], b);              // 2 register operands
        if (MODE == 2) v[c] = fma(v[c], w[c], z[c]);           // 3 register operands, same "column"
        if (MODE == 3) v[c] = fma(v[c], w[(c + 1) & 7], z[(c + 3) & 7]);   // 3 register operands, mixed
        if (MODE == 4) v[c] = fma(w[c], z[(c + 1) & 7], v[c]);  // accumulate form: acc += w*z
        if (MODE == 5) v[c] = fma(w[c & 1], z[c >> 1], v[c]);   // outer product 2 x 4: pairs share z (operand reuse cache)
        if (MODE == 6) v[c] = fma(w[(c ^ (c >> 1)) & 1], z[c >> 1], v[c]);   // serpentine: every instruction shares one operand with its predecessor
        if (MODE == 7) v[c] = fma(w[u & 7], z[c], v[c]);        // one operand fixed over 8 instructions
      }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) s += v[c] + w[c] + z[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int sms, double* buf, double* in) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2048; float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0); k<MODE><<<sms * 4, 512>>>(buf, in, iters); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
  }
  printf("%-44s %.3f ms  %.2f T DFMA/s\n", name, best, 64.0 * iters * sms * 4 * 512 / best / 1e9);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double *buf, *in; cudaMalloc(&buf, sizeof(double) * sms * 4 * 512); cudaMalloc(&in, sizeof(double) * 1024);
  double h[1024]; for (int i = 0; i < 1024; i++) h[i] = 0.5 + 1e-3 * i; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("v = fma(v, const, const)", sms, buf, in);
  run<1>("v = fma(v, w, const)", sms, buf, in);
  run<2>("v = fma(v, w, z)", sms, buf, in);
  run<3>("v = fma(v, w', z'')  (mixed registers)", sms, buf, in);
  run<4>("v = fma(w, z', v)    (accumulate)", sms, buf, in);
  run<5>("v[c] = fma(w[c&1], z[c>>1], v[c])  (pairs share z)", sms, buf, in);
  run<6>("v[c] = fma(w[serp], z[c>>1], v[c]) (serpentine)", sms, buf, in);
  run<7>("v[c] = fma(w[u], z[c], v[c])  (w fixed x8)", sms, buf, in);
  return 0;
}
