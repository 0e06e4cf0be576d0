// Microbenchmark: which part of the general-path hot loop (gen_cols, gpmpc_rollout_impl.cuh) holds the float64 pipe
// below its peak?  One CTA per SM, 2 rows x 2 columns per lane as in the kernel, column data from shared memory.
//   MODE 0  exponent FMAs + exp2s_x4 + rho                      (value mode, off-diagonal pair: 12 float64 / element)
//   MODE 1  MODE 0 with the exp table lookup replaced by a register constant (no table LDS)
//   MODE 2  MODE 0 + xi FMAs                                    (16 / element)
//   MODE 3  MODE 2 + gamma (be-weighted column sums, 8-column transpose-reduce, RED at L2): the full gradient element
//   MODE 4  only the exponent FMAs + rho (no exp)               (5 / element)
//   MODE 5  only exp2s_x4 on register inputs + rho              (8 / element)
//   MODE 6  MODE 3 with 4 rows x 1 column per lane instead of 2 x 2
// Prints clocks per float64 warp-instruction per SM sub-partition (2.0 = the DFMA pipe's peak).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../data-efficient-*/csrc gen_loop.cu -o gen_loop
#include <cstdio>
#include <cuda_runtime.h>
#include "gpmpc_common.cuh"

constexpr int EV = 4, NPTS = 512, DP = 6;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double col_reduce8(const double (&v)[8], int lane, int& col) {
  double a[4], b[2], c;
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
#pragma unroll
  for (int k = 0; k < 4; k++) { double send = u16 ? v[k] : v[k + 4]; double keep = u16 ? v[k + 4] : v[k]; a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
#pragma unroll
  for (int k = 0; k < 2; k++) { double send = u8 ? a[k] : a[k + 2]; double keep = u8 ? a[k + 2] : a[k]; b[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8); }
  { double send = u4 ? b[0] : b[1]; double keep = u4 ? b[1] : b[0]; c = keep + __shfl_xor_sync(0xffffffffu, send, 4); }
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  col = (u16 ? 4 : 0) + (u8 ? 2 : 0) + (u4 ? 1 : 0);
  return c;
}
__device__ __forceinline__ void red_add(double* addr, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory"); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, double* gam, const double* tab, int sweeps) {
  extern __shared__ __align__(16) double sm[];
  double* s_nu = sm;                       // [NPTS][DP]
  double* s_kp = s_nu + NPTS * DP;         // [NPTS]
  double* s_tabp = s_kp + NPTS;            // [2048]
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < NPTS * DP; i += blockDim.x) s_nu[i] = 0.3 * sin(0.37 * i);
  for (int i = tid; i < NPTS; i += blockDim.x) s_kp[i] = -3000.0 - 5.0 * (i % 97);
  for (int i = tid; i < EXP2S_N; i += blockDim.x) s_tabp[i] = tab[i];
  __syncthreads();
  const unsigned s_tab = exp2s_table_addr(s_tabp);
  constexpr int R = (MODE == 6) ? 4 : 2;
  double u[R][EV], rho[R], xi[R][EV], be[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    rho[r] = 0.0; be[r] = 1.0 + 0.01 * r + 1e-3 * lane;
#pragma unroll
    for (int e = 0; e < EV; e++) { u[r][e] = 100.0 * (0.5 + 0.01 * lane + 0.1 * e + r); xi[r][e] = 0.0; }
  }
  double* g_gam = gam + (size_t)blockIdx.x * NPTS;
  for (int s = 0; s < sweeps; s++) {
    const double* pn = s_nu;
    for (int j0 = 0; j0 < NPTS; j0 += 8) {
      double v[8];
      if (MODE != 6) {
#pragma unroll
        for (int jp = 0; jp < 4; jp++) {
          const int j = j0 + 2 * jp;
          double na[EV], nb[EV];
          { const double2* r2 = reinterpret_cast<const double2*>(pn); const double2 a = r2[0], b = r2[1]; na[0] = a.x; na[1] = a.y; na[2] = b.x; na[3] = b.y; }
          { const double2* r2 = reinterpret_cast<const double2*>(pn + DP); const double2 a = r2[0], b = r2[1]; nb[0] = a.x; nb[1] = a.y; nb[2] = b.x; nb[3] = b.y; }
          pn += 2 * DP;
          const double2 kk2 = *reinterpret_cast<const double2*>(s_kp + j);
          double t[4] = {kk2.x, kk2.x, kk2.y, kk2.y}, w[4];
          if (MODE != 5) {
#pragma unroll
            for (int e = 0; e < EV; e++) {
              t[0] = fma(u[0][e], na[e], t[0]);
              t[1] = fma(u[1][e], na[e], t[1]);
              t[3] = fma(u[1][e], nb[e], t[3]);
              t[2] = fma(u[0][e], nb[e], t[2]);
            }
          } else {
            t[0] += na[0]; t[1] += na[1]; t[2] += nb[0]; t[3] += nb[1];   // (4 adds: counted)
          }
          if (MODE == 4) {
#pragma unroll
            for (int q = 0; q < 4; q++) w[q] = t[q];
          } else if (MODE == 1) {
            const double SHIFT = 6755399441055744.0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
              double x = exp2s_clamp(t[q]);
              double kd = x + SHIFT;
              kd -= SHIFT;
              const double f = x - kd;
              const double tt = be[q & 1];
              double p = __fma_rn(GPMPC_EXP2S_C3, f, GPMPC_EXP2S_C2);
              p = __fma_rn(p, f, GPMPC_EXP2S_C1);
              p *= f;
              w[q] = __fma_rn(tt, p, tt);
            }
          } else {
            exp2s_x4(t, w, s_tab);
          }
          rho[0] += w[0] + w[2];
          rho[1] += w[1] + w[3];
          if (MODE == 2 || MODE == 3) {
#pragma unroll
            for (int e = 0; e < EV; e++) {
              if (e & 1) { xi[1][e] = fma(w[1], na[e], xi[1][e]); xi[0][e] = fma(w[0], na[e], xi[0][e]); }
              else       { xi[0][e] = fma(w[0], na[e], xi[0][e]); xi[1][e] = fma(w[1], na[e], xi[1][e]); }
            }
#pragma unroll
            for (int e = 0; e < EV; e++) {
              if (e & 1) { xi[1][e] = fma(w[3], nb[e], xi[1][e]); xi[0][e] = fma(w[2], nb[e], xi[0][e]); }
              else       { xi[0][e] = fma(w[2], nb[e], xi[0][e]); xi[1][e] = fma(w[3], nb[e], xi[1][e]); }
            }
          }
          if (MODE == 3) { v[2 * jp] = fma(be[0], w[0], be[1] * w[1]); v[2 * jp + 1] = fma(be[0], w[2], be[1] * w[3]); }
        }
      } else {
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
          double nj[EV];
          { const double2* r2 = reinterpret_cast<const double2*>(pn); const double2 a = r2[0], b = r2[1]; nj[0] = a.x; nj[1] = a.y; nj[2] = b.x; nj[3] = b.y; }
          pn += DP;
          const double kj = s_kp[j0 + jj];
          double t[4] = {kj, kj, kj, kj}, w[4];
#pragma unroll
          for (int e = 0; e < EV; e++) {
            t[0] = fma(u[0][e], nj[e], t[0]);
            t[1] = fma(u[1][e], nj[e], t[1]);
            t[2] = fma(u[2][e], nj[e], t[2]);
            t[3] = fma(u[3][e], nj[e], t[3]);
          }
          exp2s_x4(t, w, s_tab);
#pragma unroll
          for (int r = 0; r < 4; r++) rho[r] += w[r];
#pragma unroll
          for (int e = 0; e < EV; e++) {
            xi[0][e] = fma(w[0], nj[e], xi[0][e]);
            xi[1][e] = fma(w[1], nj[e], xi[1][e]);
            xi[2][e] = fma(w[2], nj[e], xi[2][e]);
            xi[3][e] = fma(w[3], nj[e], xi[3][e]);
          }
          v[jj] = fma(be[0], w[0], be[1] * w[1]) + fma(be[2], w[2], be[3] * w[3]);
        }
      }
      if (MODE == 3 || MODE == 6) {
        int col;
        const double tot = col_reduce8(v, lane, col);
        if ((lane & 3) == 0) red_add(g_gam + j0 + col, tot);
      }
    }
  }
  double acc = 0.0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    acc += rho[r];
#pragma unroll
    for (int e = 0; e < EV; e++) acc += xi[r][e];
  }
  out[blockIdx.x * blockDim.x + tid] = acc;
}

template <int MODE>
void run(const char* name, double per_elem, int sms, int threads, int ctas, double* out, double* gam, const double* tab, double mhz) {
  const size_t smem = sizeof(double) * (NPTS * DP + NPTS + 2048);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int sweeps = 64; float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0); k<MODE><<<sms * ctas, threads, smem>>>(out, gam, tab, sweeps); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
  }
  const int rows = (MODE == 6) ? 4 : 2;
  const double elems_per_lane = (double)sweeps * NPTS * rows;           // per thread
  const double warp_instr_per_smsp = elems_per_lane * per_elem * (threads / 32) * ctas / 4.0;
  const double clk = best * 1e-3 * mhz * 1e6;
  printf("%-58s %d x %3d thr: %7.3f ms  %.2f clk per float64 instr per SMSP (%.1f instr/element)\n", name, ctas, threads, best,
         clk / warp_instr_per_smsp, per_elem);
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(e));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount; const double mhz = p.clockRate / 1000.0;
  double *out, *gam, *tab; cudaMalloc(&out, sizeof(double) * sms * 4 * 512); cudaMalloc(&gam, sizeof(double) * sms * 4 * NPTS);
  cudaMemset(gam, 0, sizeof(double) * sms * 4 * NPTS);
  static double h[2048];
  for (int j = 0; j < 2048; j++) {
    const double v = exp2((double)j / 2048.0); unsigned long long bits; memcpy(&bits, &v, 8); bits -= (unsigned long long)j << (32 + 20 - 11); memcpy(&h[j], &bits, 8);
  }
  cudaMalloc(&tab, sizeof(h)); cudaMemcpy(tab, h, sizeof(h), cudaMemcpyHostToDevice);
  printf("clock %.0f MHz, %d SMs\n", mhz, sms);
  for (int cfg = 0; cfg < 3; cfg++) {
    const int threads = cfg == 0 ? 384 : (cfg == 1 ? 512 : 192), ctas = cfg == 2 ? 2 : 1;
    run<4>("4 exponent FMAs + rho", 5.0, sms, threads, ctas, out, gam, tab, mhz);
    run<5>("exp2s_x4 (+ 1 add) + rho", 9.0, sms, threads, ctas, out, gam, tab, mhz);
    run<1>("exponent + exp2s without the table LDS + rho", 12.0, sms, threads, ctas, out, gam, tab, mhz);
    run<0>("exponent + exp2s + rho  (value element)", 12.0, sms, threads, ctas, out, gam, tab, mhz);
    run<2>("... + xi", 16.0, sms, threads, ctas, out, gam, tab, mhz);
    run<3>("... + gamma (full gradient element, 2 rows x 2 cols)", 17.56, sms, threads, ctas, out, gam, tab, mhz);
    run<6>("full gradient element, 4 rows x 1 col", 17.28, sms, threads, ctas, out, gam, tab, mhz);
  }
  return 0;
}
