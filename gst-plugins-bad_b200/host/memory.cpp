// memory.cpp - b200vf_memory: the device-memory object the elements share.
//
// Modelled on GstCudaMemory (sys/nvcodec/gstcudamemory.c): device storage primary (:95-154), pinned staging allocated
// on the first host map (:257-300), NEED_UPLOAD / NEED_DOWNLOAD transfer flags resolved at map time (:331-407).
// New here: a memory can hold a PENDING chain of per-pixel elements instead of bytes; the chain is launched as one
// fused kernel when the bytes are needed (b200vf.h, "device memory object").
#include "memory.h"
#include <string.h>

namespace {

const size_t kSlack = 64;            // zeroed bytes after the payload, as b200vf_malloc

void drop_pending (b200vf_memory *m) {
  if (!m->pending) return;
  b200vf_memory *src = m->pending->src;
  delete m->pending;
  m->pending = nullptr;
  if (src) b200vf_memory_unref (src);
}

// A memory remembers the stream that used it last. When another stream is about to touch it, that stream is ordered
// after the last use (GstCudaMemory assumes one stream per context and synchronises; elements of one pipeline
// normally share the context's stream here too, so this is the uncommon path): caller holds m->mu.
int order_after_last_use (b200vf_memory *m, cudaStream_t s) {
  if (!m->busy || m->last_stream == s) return B200VF_OK;
  cudaEvent_t ev = nullptr;
  B200VF_CHECK_CUDA (cudaEventCreateWithFlags (&ev, cudaEventDisableTiming));
  cudaError_t e = cudaEventRecord (ev, m->last_stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent (s, ev, 0);
  cudaEventDestroy (ev);                                   // (released once the record has completed)
  if (e != cudaSuccess) {
    b200vf_set_error ("memory: ordering two streams failed: %s", cudaGetErrorString (e));
    return B200VF_E_CUDA;
  }
  return B200VF_OK;
}

// launch the pending chain of m into m->d (caller holds m->mu)
int flush_pending (b200vf_memory *m, cudaStream_t s) {
  b200vf_pending *pc = m->pending;
  if (!pc) return B200VF_OK;
  const uint8_t *d_src = nullptr;
  int rc = b200vf_memory_device_read (pc->src, s, &d_src);         // the source may itself be pending or staged on the host
  if (rc) return rc;
  if (pc->head == b200vf_pending::BAYER2RGB) {
    if (pc->has_luma || pc->has_lut)
      rc = b200vf_bayer2rgb_fused (m->ctx, d_src, pc->src_stride, pc->src_frame_stride, m->d, pc->dst_stride, pc->dst_frame_stride,
          pc->width, pc->height, pc->nframes, pc->pattern, pc->off[0], pc->off[1], pc->off[2],
          pc->has_luma ? pc->luma_table : nullptr, pc->has_lut ? pc->lut : nullptr, s);
    else
      rc = b200vf_bayer2rgb (m->ctx, d_src, pc->src_stride, pc->src_frame_stride, m->d, pc->dst_stride, pc->dst_frame_stride,
          pc->width, pc->height, pc->nframes, pc->pattern, pc->off[0], pc->off[1], pc->off[2], s);
  } else {
    rc = b200vf_lut4 (m->ctx, d_src, m->d, pc->npix, pc->lut, s);
  }
  if (rc) return rc;
  drop_pending (m);
  m->flags &= ~B200VF_MEMORY_NEED_UPLOAD;
  m->flags |= B200VF_MEMORY_NEED_DOWNLOAD;
  return B200VF_OK;
}

int upload_if_needed (b200vf_memory *m, cudaStream_t s) {
  if (!(m->flags & B200VF_MEMORY_NEED_UPLOAD)) return B200VF_OK;
  B200VF_REQUIRE (m->h, B200VF_E_INVAL, "memory: NEED_UPLOAD without a staging buffer");
  B200VF_CHECK_CUDA (cudaMemcpyAsync (m->d, m->h, m->bytes, cudaMemcpyHostToDevice, s));
  m->ctx->h2d_count.fetch_add (1, std::memory_order_relaxed);
  m->ctx->h2d_bytes.fetch_add (m->bytes, std::memory_order_relaxed);
  m->flags &= ~B200VF_MEMORY_NEED_UPLOAD;
  return B200VF_OK;
}

}  // namespace

int b200vf_memory_device_read (b200vf_memory *m, cudaStream_t s, const uint8_t **d_out) {
  B200VF_REQUIRE (m && d_out, B200VF_E_INVAL, "memory: NULL argument");
  std::lock_guard<std::mutex> g (m->mu);
  int rc = order_after_last_use (m, s);
  if (!rc) rc = flush_pending (m, s);
  if (!rc) rc = upload_if_needed (m, s);
  m->last_stream = s; m->busy = true;
  *d_out = m->d;
  return rc;
}

int b200vf_memory_device_write (b200vf_memory *m, cudaStream_t s, uint8_t **d_out) {
  B200VF_REQUIRE (m && d_out, B200VF_E_INVAL, "memory: NULL argument");
  std::lock_guard<std::mutex> g (m->mu);
  if (int rc = order_after_last_use (m, s)) return rc;
  drop_pending (m);                                      // whatever was recorded is overwritten before anyone saw it
  m->flags &= ~B200VF_MEMORY_NEED_UPLOAD;
  m->flags |= B200VF_MEMORY_NEED_DOWNLOAD;
  m->last_stream = s; m->busy = true;
  *d_out = m->d;
  return B200VF_OK;
}

int b200vf_memory_device_rw (b200vf_memory *m, cudaStream_t s, uint8_t **d_out) {
  B200VF_REQUIRE (m && d_out, B200VF_E_INVAL, "memory: NULL argument");
  std::lock_guard<std::mutex> g (m->mu);
  int rc = order_after_last_use (m, s);
  if (!rc) rc = flush_pending (m, s);
  if (!rc) rc = upload_if_needed (m, s);
  m->flags |= B200VF_MEMORY_NEED_DOWNLOAD;
  m->last_stream = s; m->busy = true;
  *d_out = m->d;
  return rc;
}

void b200vf_memory_set_pending (b200vf_memory *m, b200vf_pending *chain) {
  std::lock_guard<std::mutex> g (m->mu);
  drop_pending (m);
  m->pending = chain;
  // neither copy holds the bytes now; the flush sets NEED_DOWNLOAD
  m->flags &= ~(B200VF_MEMORY_NEED_UPLOAD | B200VF_MEMORY_NEED_DOWNLOAD);
}

B200VF_API int b200vf_memory_new (b200vf_ctx *ctx, size_t bytes, b200vf_memory **out) {
  B200VF_REQUIRE (ctx && out && bytes > 0, B200VF_E_INVAL, "memory_new: bad argument");
  *out = nullptr;
  void *d = nullptr;
  int rc = b200vf_malloc (ctx, bytes, &d);
  if (rc) return rc;
  b200vf_memory *m = new b200vf_memory ();
  m->ctx = ctx; m->device = ctx->device; m->bytes = bytes; m->d = (uint8_t *) d;
  *out = m;
  return B200VF_OK;
}

B200VF_API int b200vf_pool_acquire_memory (b200vf_pool *pool, b200vf_memory **out) {
  B200VF_REQUIRE (pool && out, B200VF_E_INVAL, "pool_acquire_memory: NULL argument");
  *out = nullptr;
  int idx = -1;
  int rc = b200vf_pool_acquire (pool, &idx);
  if (rc) return rc;
  b200vf_memory *m = new b200vf_memory ();
  b200vf_ctx *ctx = nullptr;
  // the pool's context and geometry through its public accessors
  m->d = (uint8_t *) b200vf_pool_device_ptr (pool, idx);
  m->bytes = b200vf_pool_buf_bytes (pool);
  ctx = b200vf_pool_ctx (pool);
  m->ctx = ctx; m->device = ctx->device;
  m->pool = pool; m->pool_index = idx;
  *out = m;
  return B200VF_OK;
}

B200VF_API b200vf_memory *b200vf_memory_ref (b200vf_memory *mem) {
  if (mem) mem->ref.fetch_add (1, std::memory_order_relaxed);
  return mem;
}

B200VF_API void b200vf_memory_unref (b200vf_memory *mem) {
  if (!mem) return;
  if (mem->ref.fetch_sub (1, std::memory_order_acq_rel) != 1) return;
  drop_pending (mem);                                    // releases the chain's source
  cudaSetDevice (mem->device);
  if (mem->pool) {
    // a recycled slab buffer must not be handed out while work is still queued on it
    if (mem->busy) cudaStreamSynchronize (mem->last_stream);
    b200vf_pool_release_index (mem->pool, mem->pool_index);
  } else if (mem->d) {
    cudaFree (mem->d);                                   // (synchronises)
  }
  if (mem->h) cudaFreeHost (mem->h);
  delete mem;
}

B200VF_API size_t b200vf_memory_size (const b200vf_memory *mem) { return mem ? mem->bytes : 0; }
B200VF_API unsigned b200vf_memory_flags (const b200vf_memory *mem) { return mem ? mem->flags : 0; }
B200VF_API int b200vf_memory_is_writable (const b200vf_memory *mem) { return mem && mem->ref.load () == 1; }
B200VF_API int b200vf_memory_pending_stages (const b200vf_memory *mem) { return (mem && mem->pending) ? mem->pending->stages : 0; }

B200VF_API int b200vf_memory_map (b200vf_memory *mem, int flags, void **data, void *stream) {
  B200VF_REQUIRE (mem && data && (flags & (B200VF_MAP_READ | B200VF_MAP_WRITE)), B200VF_E_INVAL, "memory_map: bad argument");
  *data = nullptr;
  cudaStream_t s = b200vf_stream (mem->ctx, stream);
  if (flags & B200VF_MAP_DEVICE) {
    uint8_t *d = nullptr;
    int rc;
    if (!(flags & B200VF_MAP_READ)) rc = b200vf_memory_device_write (mem, s, &d);
    else if (flags & B200VF_MAP_WRITE) rc = b200vf_memory_device_rw (mem, s, &d);
    else { const uint8_t *cd = nullptr; rc = b200vf_memory_device_read (mem, s, &cd); d = const_cast<uint8_t *> (cd); }
    if (rc) return rc;
    std::lock_guard<std::mutex> g (mem->mu);
    mem->map_flags = flags; mem->map_count++;
    *data = d;
    return B200VF_OK;
  }
  std::lock_guard<std::mutex> g (mem->mu);
  if (!mem->h) {                                         // gst_cuda_memory_device_memory_map, :268-300
    cudaError_t e = cudaHostAlloc ((void **) &mem->h, mem->bytes + kSlack, cudaHostAllocDefault);
    if (e != cudaSuccess) {
      b200vf_set_error ("memory_map: cudaHostAlloc(%zu): %s", mem->bytes + kSlack, cudaGetErrorString (e));
      cudaGetLastError ();
      mem->h = nullptr;
      return B200VF_E_NOMEM;
    }
    memset (mem->h + mem->bytes, 0, kSlack);
    // first host map: the device copy is the one that counts unless a writer already staged (it cannot have)
    if (!(mem->flags & B200VF_MEMORY_NEED_UPLOAD) && !mem->pending) mem->flags |= B200VF_MEMORY_NEED_DOWNLOAD;
  }
  if (flags & B200VF_MAP_READ) {
    int rc = order_after_last_use (mem, s);
    if (!rc) rc = flush_pending (mem, s);
    if (rc) return rc;
    if (mem->flags & B200VF_MEMORY_NEED_DOWNLOAD) {
      B200VF_CHECK_CUDA (cudaMemcpyAsync (mem->h, mem->d, mem->bytes, cudaMemcpyDeviceToHost, s));
      B200VF_CHECK_CUDA (cudaStreamSynchronize (s));
      mem->busy = false;                                 // s was ordered after the last use and has drained
      mem->ctx->d2h_count.fetch_add (1, std::memory_order_relaxed);
      mem->ctx->d2h_bytes.fetch_add (mem->bytes, std::memory_order_relaxed);
      mem->flags &= ~B200VF_MEMORY_NEED_DOWNLOAD;
    }
  } else {
    // write-only host map: the old contents (and anything recorded) are dead
    drop_pending (mem);
    if (mem->busy && mem->last_stream != s) B200VF_CHECK_CUDA (cudaStreamSynchronize (mem->last_stream));
    B200VF_CHECK_CUDA (cudaStreamSynchronize (s));        // device work that still reads the old bytes
    mem->busy = false;
    mem->flags &= ~B200VF_MEMORY_NEED_DOWNLOAD;
  }
  mem->map_flags = flags; mem->map_count++;
  *data = mem->h;
  return B200VF_OK;
}

B200VF_API int b200vf_memory_unmap (b200vf_memory *mem) {
  B200VF_REQUIRE (mem, B200VF_E_INVAL, "memory_unmap: NULL argument");
  std::lock_guard<std::mutex> g (mem->mu);
  B200VF_REQUIRE (mem->map_count > 0, B200VF_E_INVAL, "memory_unmap: not mapped");
  mem->map_count--;
  if ((mem->map_flags & B200VF_MAP_WRITE) && !(mem->map_flags & B200VF_MAP_DEVICE))
    mem->flags |= B200VF_MEMORY_NEED_UPLOAD;               // cuda_mem_unmap_full, :395-407
  return B200VF_OK;
}

B200VF_API int b200vf_ctx_transfer_counts (const b200vf_ctx *ctx, uint64_t *h2d_count, uint64_t *h2d_bytes, uint64_t *d2h_count,
    uint64_t *d2h_bytes)
{
  B200VF_REQUIRE (ctx, B200VF_E_INVAL, "ctx_transfer_counts: NULL argument");
  if (h2d_count) *h2d_count = ctx->h2d_count.load ();
  if (h2d_bytes) *h2d_bytes = ctx->h2d_bytes.load ();
  if (d2h_count) *d2h_count = ctx->d2h_count.load ();
  if (d2h_bytes) *d2h_bytes = ctx->d2h_bytes.load ();
  return B200VF_OK;
}
