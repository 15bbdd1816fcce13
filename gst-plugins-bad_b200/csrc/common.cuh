// common.cuh - context, error plumbing and device helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <string>
#include "../../include/b200vf.h"

#define B200VF_API extern "C" __attribute__((visibility("default")))

struct b200vf_ctx {
  int device = -1;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;
  std::atomic<uint64_t> launches{0};
  const char *last_kernel = "";
  int variant = 0;                 // 0 auto, 1 direct, 2 tma
  void *tma_encode = nullptr;      // cuTensorMapEncodeTiled entry point (driver API via cudart)
  unsigned int *tile_counters = nullptr;   // ring of work counters for dynamically scheduled kernels
  std::atomic<unsigned int> tile_counter_next{0};   // ops run on several threads
  cudaMemPool_t scratch_pool = nullptr;    // stream-ordered scratch (gaussblur pre-pass ...): never trimmed at synchronisation points
  // host <-> device transfers made by b200vf_memory (map / unmap): what a pipeline of elements that keeps its frames
  // in HBM is judged by
  std::atomic<uint64_t> h2d_count{0}, h2d_bytes{0}, d2h_count{0}, d2h_bytes{0};
  // side stream for small kernels that are independent of an op's main kernel (B200vfAux): created on first use
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
  std::mutex aux_lock;
};

// Fork / join of the context's side stream around an op's main kernel: work enqueued on stream() between construction
// and join() starts after everything that precedes the op on `s` and is waited for by everything that follows it, but
// is not ordered against what the op itself enqueues on `s` - a few small independent kernels fill the SMs that the
// main kernel's last CTAs leave idle instead of adding their launch gaps and tails to the op. The lock is held from
// fork to join: the two events are per context and ops run on several threads.
struct B200vfAux {
  b200vf_ctx *ctx;
  cudaStream_t s;
  bool active = false;
  B200vfAux (b200vf_ctx *c, cudaStream_t main_stream, bool want) : ctx (c), s (main_stream) {
    if (!want) return;
    ctx->aux_lock.lock ();
    if (!ctx->aux_stream) {
      if (cudaStreamCreateWithFlags (&ctx->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags (&ctx->aux_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags (&ctx->aux_join, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError ();
        ctx->aux_lock.unlock ();
        return;                                                   // no side stream: everything stays on `s`
      }
    }
    if (cudaEventRecord (ctx->aux_fork, s) != cudaSuccess || cudaStreamWaitEvent (ctx->aux_stream, ctx->aux_fork, 0) != cudaSuccess) {
      cudaGetLastError ();
      ctx->aux_lock.unlock ();
      return;
    }
    active = true;
  }
  cudaStream_t stream () const { return active ? ctx->aux_stream : s; }
  void join () {
    if (!active) return;
    cudaEventRecord (ctx->aux_join, ctx->aux_stream);
    cudaStreamWaitEvent (s, ctx->aux_join, 0);
    active = false;
    ctx->aux_lock.unlock ();
  }
  ~B200vfAux () { join (); }
};

void b200vf_set_error (const char *fmt, ...);

#define B200VF_CHECK_CUDA(expr)                                                      \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      b200vf_set_error ("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,          \
          cudaGetErrorString (_e));                                                  \
      return B200VF_E_CUDA;                                                          \
    }                                                                                \
  } while (0)

#define B200VF_REQUIRE(cond, status, ...)                                            \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      b200vf_set_error (__VA_ARGS__);                                                \
      return (status);                                                               \
    }                                                                                \
  } while (0)

// Every op starts here: the calling thread may be any streaming thread (a GStreamer pad task, a Python worker), whose
// current CUDA device is whatever it last used. Launches, attributes and allocations below belong to the context's
// device, so it is made current for the calling thread (left that way, as CUDA libraries do).
static inline cudaStream_t b200vf_stream (b200vf_ctx *ctx, void *stream) {
  cudaSetDevice (ctx->device);
  return stream ? (cudaStream_t) stream : ctx->stream;
}

// Opt-in to `bytes` of dynamic shared memory for kernel `fn`, once per (kernel, device): the attribute lives in the
// device's context, and ops are called from several threads (core.cu; thread-safe). Returns a b200vf status.
int b200vf_func_smem (b200vf_ctx *ctx, const void *fn, int bytes);

// Every kernel launch goes through this so that gpu_launches is a count, not a guess.
static inline int b200vf_launched (b200vf_ctx *ctx, const char *name) {
  ctx->launches.fetch_add (1, std::memory_order_relaxed);
  ctx->last_kernel = name;
  cudaError_t e = cudaGetLastError ();
  if (e != cudaSuccess) {
    b200vf_set_error ("launch of %s failed: %s", name, cudaGetErrorString (e));
    return B200VF_E_CUDA;
  }
  return B200VF_OK;
}

// host/gt_maps.cpp, for csrc/gt_device_maps.cu: the host's table entries for a list of pixels, and marble's tables
int b200vf_gt_host_index_at (const char *element, int width, int height, const char *const *prop_names, const double *prop_values,
    int nprops, int off_edge, const int32_t *pixels, size_t n, int32_t *index_out);
int b200vf_gt_marble_tables (const char *const *prop_names, const double *prop_values, int nprops, double *out2054);

// Small constant tables (LUTs, gaussian taps, colour tables) travel as
// __grid_constant__ kernel parameters: no staging buffer, nothing to
// synchronise, and the op stays asynchronous and stream-ordered.

#ifdef __CUDACC__
// ---------------------------------------------------------------- device side
__device__ __forceinline__ uint32_t ldg_u32 (const void *p) {
  return __ldg (reinterpret_cast<const unsigned int *> (p));
}
// streaming (read-once) 128-bit load / store: keep L1 for the data that is reused
__device__ __forceinline__ uint4 ld_stream_v4 (const void *p) {
  uint4 r;
  asm volatile ("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// the same for data this kernel also writes (in-place elements): no .nc, still no L1 allocation
__device__ __forceinline__ uint4 ld_na_v4 (const void *p) {
  uint4 r;
  asm volatile ("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_stream_v4 (void *p, uint4 v) {
  asm volatile ("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
      :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_v2 (void *p, uint2 v) {
  asm volatile ("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};"
      :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream_u32 (void *p, uint32_t v) {
  asm volatile ("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
// per-byte rounded-up average of 4 packed u8 = ORC avgub = (a+b+1)>>1, in 4 ops
// ((a^b) & 0xfefefefe) is one LOP3 (truth table 0x28); written as PTX because the compiler
// otherwise splits it into xor / shift / and-0x7f7f7f7f (one more ALU-pipe op per average, and this
// byte math is ALU-pipe bound: PRMT/LOP3/SHF all issue there).
__device__ __forceinline__ uint32_t avg4 (uint32_t a, uint32_t b) {
  uint32_t t;
  asm ("lop3.b32 %0, %1, %2, 0xfefefefe, 0x28;" : "=r"(t) : "r"(a), "r"(b));
  return (a | b) - (t >> 1);
}
#define PRMT(a, b, sel) __byte_perm ((a), (b), (sel))
#endif
