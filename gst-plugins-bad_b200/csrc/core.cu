// core.cu - context, error reporting, HBM buffer pool, raw memory helpers.
// Pool design follows SURVEY.md §5 / §8f-1 (precedent: sys/nvcodec/gstcudamemory.c:95-407,
// gstcudabufferpool.c:55-222): device storage primary, pinned host staging lazily.
#include "common.cuh"
#include <string.h>
#include <mutex>
#include <vector>

static thread_local char g_err[512] = "";
static const unsigned kTileCounterRing = 256;

void b200vf_set_error (const char *fmt, ...) {
  va_list ap;
  va_start (ap, fmt);
  vsnprintf (g_err, sizeof g_err, fmt, ap);
  va_end (ap);
}

B200VF_API int b200vf_version (void) { return B200VF_VERSION_MAJOR * 100 + B200VF_VERSION_MINOR; }
B200VF_API const char *b200vf_last_error (void) { return g_err; }
B200VF_API const char *b200vf_status_string (int st) {
  switch (st) {
    case B200VF_OK: return "ok";
    case B200VF_E_INVAL: return "invalid argument";
    case B200VF_E_NO_DEVICE: return "no sm_100 device (there is no CPU fallback)";
    case B200VF_E_CUDA: return "CUDA error";
    case B200VF_E_NOMEM: return "out of memory / pool drained";
    case B200VF_E_UNSUPPORTED: return "unsupported format";
    case B200VF_E_NOT_NEGOTIATED: return "not negotiated";
    case B200VF_E_NCCL: return "NCCL error";
    case B200VF_E_PROPERTY: return "bad property";
    default: return "unknown status";
  }
}

B200VF_API int b200vf_ctx_create (int device, b200vf_ctx **out) {
  B200VF_REQUIRE (out, B200VF_E_INVAL, "ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount (&n);
  if (e != cudaSuccess || n <= 0) {
    b200vf_set_error ("no CUDA device visible (%s); this library has no CPU path",
        e != cudaSuccess ? cudaGetErrorString (e) : "device count 0");
    cudaGetLastError ();
    return B200VF_E_NO_DEVICE;
  }
  B200VF_REQUIRE (device >= 0 && device < n, B200VF_E_INVAL, "ctx_create: device %d of %d", device, n);
  cudaDeviceProp prop;
  B200VF_CHECK_CUDA (cudaGetDeviceProperties (&prop, device));
  if (prop.major != 10) {
    b200vf_set_error ("device %d (%s) is sm_%d%d; kernels are built for sm_100a only",
        device, prop.name, prop.major, prop.minor);
    return B200VF_E_NO_DEVICE;
  }
  B200VF_CHECK_CUDA (cudaSetDevice (device));
  b200vf_ctx *c = new b200vf_ctx ();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  e = cudaStreamCreateWithFlags (&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    b200vf_set_error ("cudaStreamCreate: %s", cudaGetErrorString (e));
    delete c;
    return B200VF_E_CUDA;
  }
  // scratch buffers come from a private stream-ordered pool that keeps its memory: the device's default pool
  // gives everything back to the driver at each synchronisation (release threshold 0), which would re-map a
  // frame-sized scratch on every frame of a pipeline that synchronises per buffer
  {
    cudaMemPoolProps props;
    memset (&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    e = cudaMemPoolCreate (&c->scratch_pool, &props);
    if (e == cudaSuccess) {
      unsigned long long keep = ~0ull;
      e = cudaMemPoolSetAttribute (c->scratch_pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    if (e != cudaSuccess) {
      b200vf_set_error ("cudaMemPoolCreate: %s", cudaGetErrorString (e));
      if (c->scratch_pool) cudaMemPoolDestroy (c->scratch_pool);
      cudaStreamDestroy (c->stream);
      delete c;
      return B200VF_E_CUDA;
    }
  }
  e = cudaMalloc ((void **) &c->tile_counters, kTileCounterRing * sizeof (unsigned int));
  if (e != cudaSuccess) {
    b200vf_set_error ("cudaMalloc (tile counters): %s", cudaGetErrorString (e));
    cudaMemPoolDestroy (c->scratch_pool);
    cudaStreamDestroy (c->stream);
    delete c;
    return B200VF_E_NOMEM;
  }
  *out = c;
  return B200VF_OK;
}

B200VF_API void b200vf_ctx_destroy (b200vf_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice (ctx->device);
  if (ctx->stream) cudaStreamDestroy (ctx->stream);
  if (ctx->aux_stream) cudaStreamDestroy (ctx->aux_stream);
  if (ctx->aux_fork) cudaEventDestroy (ctx->aux_fork);
  if (ctx->aux_join) cudaEventDestroy (ctx->aux_join);
  if (ctx->tile_counters) cudaFree (ctx->tile_counters);
  if (ctx->scratch_pool) cudaMemPoolDestroy (ctx->scratch_pool);
  delete ctx;
}
B200VF_API int b200vf_ctx_device (const b200vf_ctx *ctx) { return ctx ? ctx->device : -1; }
B200VF_API void *b200vf_ctx_stream (const b200vf_ctx *ctx) { return ctx ? (void *) ctx->stream : nullptr; }
B200VF_API int b200vf_ctx_sm_count (const b200vf_ctx *ctx) { return ctx ? ctx->sm_count : 0; }
B200VF_API int b200vf_ctx_synchronize (b200vf_ctx *ctx, void *stream) {
  B200VF_REQUIRE (ctx, B200VF_E_INVAL, "ctx is NULL");
  B200VF_CHECK_CUDA (cudaStreamSynchronize (b200vf_stream (ctx, stream)));
  return B200VF_OK;
}
B200VF_API uint64_t b200vf_ctx_launch_count (const b200vf_ctx *ctx) { return ctx ? ctx->launches.load () : 0; }
B200VF_API const char *b200vf_ctx_last_kernel (const b200vf_ctx *ctx) { return ctx ? ctx->last_kernel : ""; }
B200VF_API int b200vf_ctx_set_variant (b200vf_ctx *ctx, int variant) {
  B200VF_REQUIRE (ctx && variant >= 0 && variant <= 2, B200VF_E_INVAL, "set_variant: %d", variant);
  ctx->variant = variant;
  return B200VF_OK;
}

// ----------------------------------------------------------------- raw memory
static const size_t kSlack = 64;   // zeroed bytes after every buffer

B200VF_API int b200vf_malloc (b200vf_ctx *ctx, size_t bytes, void **d_out) {
  B200VF_REQUIRE (ctx && d_out, B200VF_E_INVAL, "malloc: NULL argument");
  B200VF_CHECK_CUDA (cudaSetDevice (ctx->device));
  void *p = nullptr;
  cudaError_t e = cudaMalloc (&p, bytes + kSlack);
  if (e != cudaSuccess) {
    b200vf_set_error ("cudaMalloc(%zu): %s", bytes + kSlack, cudaGetErrorString (e));
    cudaGetLastError ();
    return B200VF_E_NOMEM;
  }
  // zero on the context stream and wait: the legacy default stream does not order
  // against non-blocking streams, so a plain cudaMemset could land after a later kernel
  B200VF_CHECK_CUDA (cudaMemsetAsync (p, 0, bytes + kSlack, ctx->stream));
  B200VF_CHECK_CUDA (cudaStreamSynchronize (ctx->stream));
  *d_out = p;
  return B200VF_OK;
}
B200VF_API int b200vf_free (b200vf_ctx *ctx, void *d_ptr) {
  B200VF_REQUIRE (ctx, B200VF_E_INVAL, "free: ctx is NULL");
  if (d_ptr) B200VF_CHECK_CUDA (cudaFree (d_ptr));
  return B200VF_OK;
}
B200VF_API int b200vf_host_alloc (size_t bytes, void **h_out) {
  B200VF_REQUIRE (h_out, B200VF_E_INVAL, "host_alloc: NULL argument");
  cudaError_t e = cudaHostAlloc (h_out, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    b200vf_set_error ("cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString (e));
    cudaGetLastError ();
    return B200VF_E_NOMEM;
  }
  return B200VF_OK;
}
B200VF_API int b200vf_host_free (void *h_ptr) {
  if (h_ptr) B200VF_CHECK_CUDA (cudaFreeHost (h_ptr));
  return B200VF_OK;
}
B200VF_API int b200vf_memcpy_h2d (b200vf_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, void *stream) {
  B200VF_REQUIRE (ctx && d_dst && h_src, B200VF_E_INVAL, "memcpy_h2d: NULL argument");
  B200VF_CHECK_CUDA (cudaMemcpyAsync (d_dst, h_src, bytes, cudaMemcpyHostToDevice, b200vf_stream (ctx, stream)));
  return B200VF_OK;
}
B200VF_API int b200vf_memcpy_d2h (b200vf_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, void *stream) {
  B200VF_REQUIRE (ctx && h_dst && d_src, B200VF_E_INVAL, "memcpy_d2h: NULL argument");
  B200VF_CHECK_CUDA (cudaMemcpyAsync (h_dst, d_src, bytes, cudaMemcpyDeviceToHost, b200vf_stream (ctx, stream)));
  return B200VF_OK;
}

// ----------------------------------------------------------------------- pool
struct b200vf_pool {
  b200vf_ctx *ctx = nullptr;
  int device = -1;                 // cached: destroy must not read a context that was destroyed first
  size_t buf_bytes = 0, pitch = 0;
  int n = 0;
  uint8_t *d_base = nullptr;       // one slab: n buffers at a fixed pitch (batch ops index it directly)
  uint8_t *h_base = nullptr;       // pinned staging slab, allocated on first host map
  std::vector<int> free_list;
  std::mutex mu;
};

B200VF_API int b200vf_pool_create (b200vf_ctx *ctx, size_t buf_bytes, int n_bufs, b200vf_pool **out) {
  B200VF_REQUIRE (ctx && out && buf_bytes > 0 && n_bufs > 0, B200VF_E_INVAL, "pool_create: bad argument");
  b200vf_pool *p = new b200vf_pool ();
  p->ctx = ctx;
  p->device = ctx->device;
  p->buf_bytes = buf_bytes;
  p->pitch = (buf_bytes + kSlack + 255) & ~(size_t) 255;   // 256 B aligned: TMA bases, 128-bit vectors
  p->n = n_bufs;
  cudaSetDevice (ctx->device);
  cudaError_t e = cudaMalloc ((void **) &p->d_base, p->pitch * n_bufs);
  if (e != cudaSuccess) {
    b200vf_set_error ("pool_create: cudaMalloc(%zu): %s", p->pitch * n_bufs, cudaGetErrorString (e));
    cudaGetLastError ();
    delete p;
    return B200VF_E_NOMEM;
  }
  cudaMemsetAsync (p->d_base, 0, p->pitch * n_bufs, ctx->stream);
  cudaStreamSynchronize (ctx->stream);
  for (int i = n_bufs - 1; i >= 0; i--) p->free_list.push_back (i);
  *out = p;
  return B200VF_OK;
}
B200VF_API void b200vf_pool_destroy (b200vf_pool *pool) {
  if (!pool) return;
  cudaSetDevice (pool->device);
  if (pool->d_base) cudaFree (pool->d_base);
  if (pool->h_base) cudaFreeHost (pool->h_base);
  delete pool;
}
B200VF_API int b200vf_pool_acquire (b200vf_pool *pool, int *buf_index) {
  B200VF_REQUIRE (pool && buf_index, B200VF_E_INVAL, "pool_acquire: NULL argument");
  std::lock_guard<std::mutex> g (pool->mu);
  if (pool->free_list.empty ()) {
    b200vf_set_error ("pool_acquire: all %d buffers are out", pool->n);
    return B200VF_E_NOMEM;
  }
  *buf_index = pool->free_list.back ();
  pool->free_list.pop_back ();
  return B200VF_OK;
}
B200VF_API int b200vf_pool_release (b200vf_pool *pool, int buf_index) {
  B200VF_REQUIRE (pool && buf_index >= 0 && buf_index < pool->n, B200VF_E_INVAL, "pool_release: index %d", buf_index);
  std::lock_guard<std::mutex> g (pool->mu);
  for (int i : pool->free_list)
    B200VF_REQUIRE (i != buf_index, B200VF_E_INVAL, "pool_release: buffer %d released twice", buf_index);
  pool->free_list.push_back (buf_index);
  return B200VF_OK;
}
B200VF_API void *b200vf_pool_device_ptr (b200vf_pool *pool, int i) {
  return (pool && i >= 0 && i < pool->n) ? pool->d_base + pool->pitch * i : nullptr;
}
B200VF_API void *b200vf_pool_host_ptr (b200vf_pool *pool, int i) {
  if (!pool || i < 0 || i >= pool->n) return nullptr;
  std::lock_guard<std::mutex> g (pool->mu);
  if (!pool->h_base) {
    if (cudaHostAlloc ((void **) &pool->h_base, pool->pitch * pool->n, cudaHostAllocDefault) != cudaSuccess) {
      b200vf_set_error ("pool_host_ptr: cudaHostAlloc(%zu) failed", pool->pitch * pool->n);
      cudaGetLastError ();
      pool->h_base = nullptr;
      return nullptr;
    }
    memset (pool->h_base, 0, pool->pitch * pool->n);
  }
  return pool->h_base + pool->pitch * i;
}
// hooks for b200vf_memory (host/memory.cpp)
int b200vf_pool_release_index (b200vf_pool *pool, int index) { return b200vf_pool_release (pool, index); }
b200vf_ctx *b200vf_pool_ctx (b200vf_pool *pool) { return pool ? pool->ctx : nullptr; }
B200VF_API size_t b200vf_pool_buf_bytes (const b200vf_pool *pool) { return pool ? pool->buf_bytes : 0; }
B200VF_API size_t b200vf_pool_buf_pitch (const b200vf_pool *pool) { return pool ? pool->pitch : 0; }
B200VF_API int b200vf_pool_upload (b200vf_pool *pool, int i, const void *host_src, size_t bytes, void *stream) {
  B200VF_REQUIRE (pool && host_src && i >= 0 && i < pool->n && bytes <= pool->buf_bytes, B200VF_E_INVAL,
      "pool_upload: bad argument");
  B200VF_CHECK_CUDA (cudaMemcpyAsync (pool->d_base + pool->pitch * i, host_src, bytes, cudaMemcpyHostToDevice,
      b200vf_stream (pool->ctx, stream)));
  return B200VF_OK;
}
B200VF_API int b200vf_pool_download (b200vf_pool *pool, int i, void *host_dst, size_t bytes, void *stream) {
  B200VF_REQUIRE (pool && host_dst && i >= 0 && i < pool->n && bytes <= pool->buf_bytes, B200VF_E_INVAL,
      "pool_download: bad argument");
  B200VF_CHECK_CUDA (cudaMemcpyAsync (host_dst, pool->d_base + pool->pitch * i, bytes, cudaMemcpyDeviceToHost,
      b200vf_stream (pool->ctx, stream)));
  return B200VF_OK;
}

// Work counters for dynamically scheduled persistent kernels: a ring of 256 counters (allocated with the context), each zeroed
// on the launching stream right before its launch (stream-ordered, so launches on the same stream
// never share a live counter; 256 in-flight launches across streams would be needed to collide).
int b200vf_next_tile_counter (b200vf_ctx *ctx, cudaStream_t s, unsigned int **out) {
  unsigned int *c = ctx->tile_counters + (ctx->tile_counter_next.fetch_add (1, std::memory_order_relaxed) % kTileCounterRing);
  B200VF_CHECK_CUDA (cudaMemsetAsync (c, 0, sizeof (unsigned int), s));
  *out = c;
  return B200VF_OK;
}

// ---------------------------------------------------------------- per-device kernel attributes
#include <mutex>
#include <set>
#include <utility>
int b200vf_func_smem (b200vf_ctx *ctx, const void *fn, int bytes)
{
  static std::mutex mu;
  static std::set<std::pair<const void *, int>> done;      // (kernel, device) pairs already opted in
  std::lock_guard<std::mutex> lock (mu);
  const std::pair<const void *, int> key (fn, ctx->device);
  if (done.count (key)) return B200VF_OK;
  B200VF_CHECK_CUDA (cudaSetDevice (ctx->device));
  B200VF_CHECK_CUDA (cudaFuncSetAttribute (fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.insert (key);
  return B200VF_OK;
}
