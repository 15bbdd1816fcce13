/* gstb200vfmemory.h - GstAllocator / GstBufferPool over b200vf_memory (the HBM-backed pool of SURVEY.md 8f rank 1).
 *
 * Precedent in the reference tree: sys/nvcodec/gstcudamemory.[ch] (allocator, GST_MAP_CUDA, transfer flags) and
 * sys/nvcodec/gstcudabufferpool.c:55-222 (pool handing out such memories with a GstVideoMeta).
 */
#ifndef GST_B200VF_MEMORY_H
#define GST_B200VF_MEMORY_H

#include <gst/gst.h>
#include <gst/video/video.h>
#include "b200vf.h"

G_BEGIN_DECLS

#define GST_B200VF_MEMORY_TYPE "B200VFMemory"
#define GST_CAPS_FEATURE_MEMORY_B200VF "memory:B200VFMemory"
/* map flag asking for the HBM address instead of the pinned staging copy (GST_MAP_CUDA, gstcudamemory.h:60) */
#define GST_MAP_B200VF (GST_MAP_FLAG_LAST << 1)

typedef struct _GstB200vfMemory
{
  GstMemory mem;
  b200vf_memory *vf;            /* owned reference */
} GstB200vfMemory;

GstAllocator *gst_b200vf_allocator_new (b200vf_ctx * ctx);
/* a fresh device memory of `size` bytes, or one wrapping `vf` (takes the reference) */
GstMemory *gst_b200vf_allocator_alloc (GstAllocator * allocator, gsize size);
GstMemory *gst_b200vf_allocator_wrap (GstAllocator * allocator, b200vf_memory * vf);
gboolean gst_is_b200vf_memory (GstMemory * mem);
b200vf_memory *gst_b200vf_memory_peek (GstMemory * mem);
/* the b200vf_memory behind a buffer made of exactly one such GstMemory, else NULL */
b200vf_memory *gst_b200vf_buffer_peek (GstBuffer * buffer);

GstBufferPool *gst_b200vf_buffer_pool_new (b200vf_ctx * ctx);

G_END_DECLS
#endif
