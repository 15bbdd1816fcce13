// Developer probe: FMA-pipe throughput of the gaussblur tap pattern (8 outputs x 4 channels per thread, rotating
// 8-sample window), packed (FMUL2 + FFMA2 by an opaque 1.0) against scalar (FMUL + FADD), and the FMA-only forms.
// Prints lane-operations per clock per SMSP (32 = the FP32 pipe's peak). nvcc -arch=sm_100a -fmad=false -O3.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
typedef unsigned long long f32x2;
struct px4 { f32x2 lo, hi; };
template <int MODE> __global__ void __launch_bounds__ (256, 2) k (float4 *out, const float4 *in, f32x2 one, long long *cyc, int iters) {
  __shared__ f32x2 s_k2[32];
  if (threadIdx.x < 32) s_k2[threadIdx.x] = 0x3f0000003f000000ull + threadIdx.x * 0x0000100000001000ull;
  __syncthreads ();
  px4 W[8], acc[8];
  float w[8][4], a[8][4];
  for (int j = 0; j < 8; j++) {
    float4 v = in[threadIdx.x * 8 + j];
    w[j][0] = v.x; w[j][1] = v.y; w[j][2] = v.z; w[j][3] = v.w;
    asm ("mov.b64 %0, {%1, %2};" : "=l"(W[j].lo) : "f"(v.x), "f"(v.y)); asm ("mov.b64 %0, {%1, %2};" : "=l"(W[j].hi) : "f"(v.z), "f"(v.w));
    acc[j].lo = acc[j].hi = 0; a[j][0] = a[j][1] = a[j][2] = a[j][3] = 0.f;
  }
  long long t0 = clock64 ();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
      const f32x2 coef = s_k2[(it & 3) * 8 + kk];
      float cf; asm ("{ .reg .b32 t; mov.b64 {%0, t}, %1; }" : "=f"(cf) : "l"(coef));
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int s = (j + kk) & 7;
        if (MODE == 0) {
          f32x2 m0, m1;
          asm volatile ("mul.rn.f32x2 %0, %1, %2;" : "=l"(m0) : "l"(W[s].lo), "l"(coef));
          asm volatile ("mul.rn.f32x2 %0, %1, %2;" : "=l"(m1) : "l"(W[s].hi), "l"(coef));
          asm volatile ("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j].lo) : "l"(m0), "l"(one));
          asm volatile ("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j].hi) : "l"(m1), "l"(one));
        } else if (MODE == 1) {
#pragma unroll
          for (int c = 0; c < 4; c++) a[j][c] = __fadd_rn (a[j][c], __fmul_rn (w[s][c], cf));
        } else if (MODE == 2) {
          asm volatile ("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j].lo) : "l"(W[s].lo), "l"(coef));
          asm volatile ("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j].hi) : "l"(W[s].hi), "l"(coef));
        } else {
#pragma unroll
          for (int c = 0; c < 4; c++) a[j][c] = __fmaf_rn (w[s][c], cf, a[j][c]);
        }
      }
    }
  }
  long long t1 = clock64 ();
  for (int j = 0; j < 8; j++) {
    float x, y, z, u;
    asm ("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[j].lo)); asm ("mov.b64 {%0, %1}, %2;" : "=f"(z), "=f"(u) : "l"(acc[j].hi));
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 8 + j] = make_float4 (x + a[j][0], y + a[j][1], z + a[j][2], u + a[j][3]);
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run (const char *name, int ctas_per_sm, int threads) {
  float4 *out, *in; long long *cyc;
  cudaMalloc (&out, (size_t) 148 * 2 * 256 * 8 * 16); cudaMalloc (&in, 256 * 8 * 16); cudaMemset (in, 0x3c, 256 * 8 * 16); cudaMalloc (&cyc, 8);
  const int iters = 2048;
  k<MODE><<<148 * ctas_per_sm, threads>>> (out, in, 0x3f8000003f800000ull, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize ();
  long long c; cudaMemcpy (&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double lane_ops_per_thread = (double) iters * 8 * 8 * 4 * ((MODE < 2) ? 2 : 1);        // mul + add, or one fma
  const double warps_per_smsp = ctas_per_sm * threads / 32 / 4.0;
  printf ("%-22s %d CTA/SM x %3d thr: %6.2f lane-ops/clk/SMSP of 32  (%s)\n", name, ctas_per_sm, threads,
      lane_ops_per_thread * 32 * warps_per_smsp / c, cudaGetErrorString (e));
  cudaFree (out); cudaFree (in); cudaFree (cyc);
}
int main () {
  for (int ctas : {1, 2}) for (int thr : {128, 256}) {
    run<0> ("packed FMUL2+FFMA2", ctas, thr); run<1> ("scalar FMUL+FADD", ctas, thr);
    run<2> ("packed FFMA2", ctas, thr); run<3> ("scalar FFMA", ctas, thr);
  }
  return 0;
}
