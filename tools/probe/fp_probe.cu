// Developer probe: issue rate of packed f32x2 vs scalar fp32 ops on this part (warp-instructions per clock per SM).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
typedef unsigned long long f32x2;
template <int MODE> __global__ void k(f32x2* out, f32x2 a0, f32x2 b0, long long* cyc, int iters) {
  f32x2 acc[16]; float fa[16];
  for (int i = 0; i < 16; i++) { acc[i] = a0 + i; fa[i] = (float) i; }
  float fb = __uint_as_float((unsigned) b0);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (MODE == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(acc[i]) : "l"(b0));
      if (MODE == 1) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(acc[i]) : "l"(b0));
      if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(fa[i]) : "f"(fb));
      if (MODE == 3) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(fa[i]) : "f"(fb));
      if (MODE == 4) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(fa[i]) : "f"(fb));
    }
  }
  long long t1 = clock64();
  f32x2 s = 0; for (int i = 0; i < 16; i++) s += acc[i] + (f32x2) fa[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int warps_per_sm) {
  f32x2* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  int iters = 4096, threads = warps_per_sm * 32;
  k<MODE><<<148, threads>>>(out, 0x3f8000003f800000ull, 0x3f8000013f800001ull, cyc, iters);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double instr = (double) iters * 16 * warps_per_sm;   // warp-instructions per SM
  printf("%-12s warps/SM=%2d: %.3f warp-instr/clk/SM (%.2f per SMSP)\n", name, warps_per_sm, instr / c, instr / c / 4);
}
int main() {
  for (int w : {4, 8, 16, 32}) { run<0>("FFMA2", w); run<1>("FMUL2", w); run<2>("FFMA", w); run<3>("FMUL", w); run<4>("FADD", w); }
  return 0;
}
