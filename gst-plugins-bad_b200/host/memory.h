// memory.h - internals of b200vf_memory shared by memory.cpp and elements.cpp.
#pragma once
#include "../csrc/common.cuh"
#include <atomic>
#include <mutex>

// A chain of per-pixel operations recorded on a memory instead of being launched (see memory.cpp): flushed as ONE
// kernel when the memory's bytes are needed.
struct b200vf_pending {
  enum Head { BAYER2RGB, LUT_ONLY } head = LUT_ONLY;
  b200vf_memory *src = nullptr;        // retained until the chain is flushed or dropped
  // BAYER2RGB head (b200vf_bayer2rgb's arguments)
  int width = 0, height = 0, nframes = 1, src_stride = 0, dst_stride = 0, pattern = 0, off[3] = { 0, 0, 0 };
  size_t src_frame_stride = 0, dst_frame_stride = 0;
  // LUT_ONLY head
  size_t npix = 0;
  // fused stages, in order: one luma-mapped coloreffects preset (BAYER2RGB head only), then per-byte-position LUTs
  bool has_luma = false, has_lut = false;
  uint8_t luma_table[768];
  uint8_t lut[4][256];
  int stages = 1;                      // elements folded into the chain (introspection / tests)
};

struct b200vf_memory {
  b200vf_ctx *ctx = nullptr;
  int device = -1;
  size_t bytes = 0;
  uint8_t *d = nullptr;                // device storage (primary)
  uint8_t *h = nullptr;                // pinned staging, allocated on the first host map
  unsigned flags = 0;                  // B200VF_MEMORY_NEED_UPLOAD / NEED_DOWNLOAD
  std::atomic<int> ref{1};
  b200vf_pool *pool = nullptr;         // owner pool (device storage is a slab buffer) or NULL (own allocation)
  int pool_index = -1;
  b200vf_pending *pending = nullptr;
  int map_flags = 0, map_count = 0;
  cudaStream_t last_stream = nullptr;  // stream of the last device access; `busy`: work may still be queued on it
  bool busy = false;
  std::mutex mu;
};

// device pointer valid for reading on `stream`: flushes a pending chain, uploads staged host bytes
int b200vf_memory_device_read (b200vf_memory *m, cudaStream_t s, const uint8_t **d_out);
// device pointer that is about to be overwritten on `stream`: drops any pending chain, marks the staging copy stale
int b200vf_memory_device_write (b200vf_memory *m, cudaStream_t s, uint8_t **d_out);
// the same for read-modify-write (in-place elements)
int b200vf_memory_device_rw (b200vf_memory *m, cudaStream_t s, uint8_t **d_out);
void b200vf_memory_set_pending (b200vf_memory *m, b200vf_pending *chain);      // takes ownership; drops an old chain
// pool hooks (core.cu)
int b200vf_pool_release_index (b200vf_pool *pool, int index);
b200vf_ctx *b200vf_pool_ctx (b200vf_pool *pool);
