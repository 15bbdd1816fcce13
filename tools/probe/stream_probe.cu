// Developer probe: the streaming form of the gaussblur tap loop. A thread owns one (column, channel pair) and
// walks down the rows with a register ring of 2C+1 packed accumulators; the taps are bitwise symmetric
// (k[i] == k[2C-i]), so the product of a sample with tap k serves two outputs: C+1 FMUL2 + 2C+1 FFMA2 (by an
// opaque 1.0: separately rounded add) per sample instead of 2 x (2C+1).
// Prints packed instructions per clock per SMSP (0.5 = the FP32 pipe's peak for f32x2) and the equivalent
// "unshared" lane-operations per clock per SMSP the same outputs would have cost (of 32).
// nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o stream_probe stream_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 mul2 (f32x2 a, f32x2 b) { f32x2 r; asm volatile ("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ void acc2 (f32x2 &acc, f32x2 m, f32x2 one) { asm volatile ("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(m), "l"(one)); }
__device__ __forceinline__ f32x2 fma2 (f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm ("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mulp (f32x2 a, f32x2 b) { f32x2 r; asm ("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2_rm (f32x2 a, f32x2 b) { f32x2 r; asm ("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 div2 (f32x2 a, f32x2 nb2, f32x2 rb2) {
  f32x2 q = mulp (a, rb2);
  f32x2 r = fma2 (nb2, q, a);
  q = fma2 (r, rb2, q);
  r = fma2 (nb2, q, a);
  return fma2 (r, rb2, q);
}

struct Taps { f32x2 k[16]; };

template <int C, int I, int EPI>
__device__ __forceinline__ void vstep (f32x2 (&A)[2 * C + 1], const f32x2 v, const f32x2 (&K)[C + 1], const f32x2 one,
    const f32x2 nb2, const f32x2 rb2, uint32_t *outp, int lane)
{
  constexpr int R = 2 * C + 1;
  f32x2 m[C + 1];
#pragma unroll
  for (int k = 0; k <= C; k++) m[k] = mul2 (v, K[k]);
  A[(I + C) % R] = m[0];
#pragma unroll
  for (int k = 1; k <= C; k++) acc2 (A[(I + C - k) % R], m[k], one);
#pragma unroll
  for (int k = C - 1; k >= 0; k--) acc2 (A[(I - C + k + R) % R], m[k], one);
  if (EPI) {
    const f32x2 q = div2 (A[(I + C + 1) % R], nb2, rb2);
    const f32x2 H2 = 0x3F0000003F000000ull, M2 = 0x4B0000004B000000ull;
    const f32x2 y = add2_rm (add2_rm (q, H2), M2);
    uint32_t lo = (uint32_t) y, hi = (uint32_t) (y >> 32);
    uint32_t two = __byte_perm (lo, hi, 0x0040);              // 2 bytes
    uint32_t other = __shfl_xor_sync (0xffffffffu, two, 1);
    uint32_t word = __byte_perm (two, other, 0x5410);
    if (!(lane & 1)) outp[I * 64] = word;
  }
}

template <int C, int I, int EPI> struct Unroll {
  static __device__ __forceinline__ void run (f32x2 (&A)[2 * C + 1], const f32x2 *col, const f32x2 (&K)[C + 1], const f32x2 one,
      const f32x2 nb2, const f32x2 rb2, uint32_t *outp, int lane) {
    Unroll<C, I - 1, EPI>::run (A, col, K, one, nb2, rb2, outp, lane);
    vstep<C, I, EPI> (A, col[I * 128], K, one, nb2, rb2, outp, lane);
  }
};
template <int C, int EPI> struct Unroll<C, -1, EPI> {
  static __device__ __forceinline__ void run (f32x2 (&)[2 * C + 1], const f32x2 *, const f32x2 (&)[C + 1], const f32x2,
      const f32x2, const f32x2, uint32_t *, int) {}
};

template <int C, int EPI, int THREADS>
__global__ void __launch_bounds__ (THREADS, 1) k (uint32_t *out, const __grid_constant__ Taps taps, f32x2 one, float b, long long *cyc, int iters) {
  constexpr int R = 2 * C + 1;
  extern __shared__ f32x2 tile[];                              // [R rows][128 (column, pair)] f32x2, reused every iteration
  for (int i = threadIdx.x; i < R * 128; i += THREADS) tile[i] = 0x3f8100003f800000ull + (uint64_t) i * 0x100000001ull;
  __syncthreads ();
  f32x2 K[C + 1], A[R];
#pragma unroll
  for (int i = 0; i <= C; i++) K[i] = taps.k[i];
#pragma unroll
  for (int i = 0; i < R; i++) A[i] = 0;
  const float rb = 1.f / b;
  f32x2 nb2, rb2;
  asm ("mov.b64 %0, {%1, %1};" : "=l"(nb2) : "f"(-b));
  asm ("mov.b64 %0, {%1, %1};" : "=l"(rb2) : "f"(rb));
  const f32x2 *col = tile + (threadIdx.x & 127);
  uint32_t *outp = out + (size_t) blockIdx.x * THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31;
  long long t0 = clock64 ();
#pragma unroll 1
  for (int it = 0; it < iters; it++) Unroll<C, R - 1, EPI>::run (A, col, K, one, nb2, rb2, outp, lane);
  long long t1 = clock64 ();
  f32x2 s = 0;
#pragma unroll
  for (int i = 0; i < R; i++) s ^= A[i];
  out[(size_t) blockIdx.x * THREADS + threadIdx.x] = (uint32_t) s ^ (uint32_t) (s >> 32);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int C, int EPI, int THREADS> void run () {
  uint32_t *out; long long *cyc;
  cudaMalloc (&out, (size_t) 148 * THREADS * 4 * 64); cudaMalloc (&cyc, 8);
  Taps t;
  for (int i = 0; i < 16; i++) { float f = 0.01f + 0.003f * i; uint32_t u; memcpy (&u, &f, 4); t.k[i] = ((uint64_t) u << 32) | u; }
  const int iters = 512, R = 2 * C + 1;
  const int smem = R * 128 * 8;
  cudaFuncSetAttribute (k<C, EPI, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; rep++) k<C, EPI, THREADS><<<148, THREADS, smem>>> (out, t, 0x3f8000003f800000ull, 0.99999994f, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize ();
  long long c; cudaMemcpy (&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double samples = (double) iters * R;
  const double packed = samples * (C + 1 + 2 * C);                  // FMUL2 + FFMA2 actually issued (tap work only)
  const double warps_per_smsp = THREADS / 32 / 4.0;
  const double unshared_lane_ops = samples * 2 * R * 2 * 32;        // mul + add per tap, 2 lanes per packed op, 32 lanes
  printf ("C=%2d epi=%d %3d thr/SM: %.3f packed tap instr/clk/SMSP (peak 0.5), equivalent unshared lane-ops/clk/SMSP %.1f of 32  (%s)\n",
      C, EPI, THREADS, packed * warps_per_smsp / c, unshared_lane_ops * warps_per_smsp / c, cudaGetErrorString (e));
  cudaFree (out); cudaFree (cyc);
}
int main () {
  run<13, 0, 256> (); run<13, 0, 512> (); run<13, 1, 256> (); run<13, 1, 512> ();
  run<4, 0, 512> (); run<4, 1, 512> (); run<4, 1, 1024> ();
  return 0;
}
