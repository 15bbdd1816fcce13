// Developer probe: which TMA box shapes does a plain 3-D u32 tensor accept on this part?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
__global__ void k(const __grid_constant__ CUtensorMap m, int bytes, int c0, int c1, uint32_t* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :: "r"(d), "l"(&m), "r"(b), "r"(c0), "r"(c1), "r"(0) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" :: "r"(b) : "memory");
  if (threadIdx.x == 0) out[0] = ((uint32_t*)sm)[0] + ((uint32_t*)sm)[bytes/4-1];
}
int main(int argc, char** argv) {
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q); Enc enc = (Enc)fn;
  int W = atoi(argv[1]), H = atoi(argv[2]), bx = atoi(argv[3]), by = atoi(argv[4]), c0 = atoi(argv[5]), c1 = atoi(argv[6]), extra = argc > 7 ? atoi(argv[7]) : 256;
  uint32_t* buf; cudaMalloc(&buf, (size_t)W*H*4*2); cudaMemset(buf, 1, (size_t)W*H*4*2);
  uint32_t* out; cudaMalloc(&out, 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024);
  CUtensorMap m; cuuint64_t gd[3] = {(cuuint64_t)W, (cuuint64_t)H, 2}; cuuint64_t gs[2] = {(cuuint64_t)W*4, (cuuint64_t)W*H*4}; cuuint32_t b[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1}; cuuint32_t es[3] = {1,1,1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, buf, gd, gs, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  int bytes = bx*by*4;
  k<<<1, 64, bytes + extra>>>(m, bytes, c0, c1, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("W=%d H=%d box %3dx%2d (%6d B, smem %d) coords %d,%d: encode=%d run=%s\n", W, H, bx, by, bytes, bytes + extra, c0, c1, (int)r, cudaGetErrorString(e));
  return 0;
}
