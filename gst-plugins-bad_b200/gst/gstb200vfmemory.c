/* gstb200vfmemory.c - GstAllocator and GstBufferPool of the HBM-backed pool.
 *
 * GstB200vfAllocator wraps b200vf_memory (device storage primary, pinned staging on the first host map, transfer
 * flags; host/memory.cpp) as GstMemory the way GstCudaAllocator wraps CUdeviceptr (sys/nvcodec/gstcudamemory.c:
 * 95-154 alloc, 331-407 map / unmap): gst_memory_map (GST_MAP_READ | GST_MAP_B200VF) yields the HBM address,
 * a plain gst_memory_map the pinned staging copy (downloaded when stale), so any sysmem element downstream keeps
 * working. GstB200vfBufferPool hands out buffers of one such memory plus a GstVideoMeta
 * (sys/nvcodec/gstcudabufferpool.c:55-222); its memories come from one b200vf_pool slab (256-byte aligned buffer
 * pitch, >= 64 zeroed slack bytes: the batch layout the kernels take).
 *
 * Compile-checked against tests/stubs/ in CI (tests/test_shells_cpu.py); built for real by gst/meson.build.
 */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#include <string.h>
#include "gstb200vfmemory.h"

GST_DEBUG_CATEGORY_STATIC (b200vf_memory_debug);
#define GST_CAT_DEFAULT b200vf_memory_debug

/* ------------------------------------------------------------------ allocator */
typedef struct
{
  GstAllocator parent;
  b200vf_ctx *ctx;              /* not owned: the element that created the allocator owns the context */
} GstB200vfAllocator;

typedef struct
{
  GstAllocatorClass parent_class;
} GstB200vfAllocatorClass;

GType gst_b200vf_allocator_get_type (void);
G_DEFINE_TYPE (GstB200vfAllocator, gst_b200vf_allocator, GST_TYPE_ALLOCATOR);

static GstMemory *
b200vf_allocator_dummy_alloc (GstAllocator * allocator, gsize size, GstAllocationParams * params)
{
  /* gst_allocator_alloc () carries no device context: use gst_b200vf_allocator_alloc (gstcudamemory.c:86-93 does the same) */
  g_return_val_if_reached (NULL);
}

static void
b200vf_allocator_free (GstAllocator * allocator, GstMemory * memory)
{
  GstB200vfMemory *m = (GstB200vfMemory *) memory;
  b200vf_memory_unref (m->vf);  /* a pool memory goes back to its slab */
  g_free (m);
}

static gpointer
b200vf_mem_map (GstMemory * memory, gsize maxsize, GstMapFlags flags)
{
  GstB200vfMemory *m = (GstB200vfMemory *) memory;
  int f = 0;
  void *data = NULL;
  if (flags & GST_MAP_READ) f |= B200VF_MAP_READ;
  if (flags & GST_MAP_WRITE) f |= B200VF_MAP_WRITE;
  if (flags & GST_MAP_B200VF) f |= B200VF_MAP_DEVICE;
  if (b200vf_memory_map (m->vf, f, &data, NULL) != B200VF_OK) {
    GST_ERROR ("map failed: %s", b200vf_last_error ());
    return NULL;
  }
  return data;
}

static void
b200vf_mem_unmap (GstMemory * memory)
{
  b200vf_memory_unmap (((GstB200vfMemory *) memory)->vf);
}

static GstMemory *
b200vf_mem_copy (GstMemory * memory, gssize offset, gssize size)
{
  /* rare (gst_buffer_copy_deep): through the staging copies */
  GstB200vfMemory *src = (GstB200vfMemory *) memory;
  GstMemory *copy;
  GstMapInfo in, out;
  if (size == -1) size = (gssize) memory->size - offset;
  copy = gst_b200vf_allocator_alloc (memory->allocator, b200vf_memory_size (src->vf));
  if (!copy) return NULL;
  if (!gst_memory_map (memory, &in, GST_MAP_READ)) {
    gst_memory_unref (copy);
    return NULL;
  }
  if (!gst_memory_map (copy, &out, GST_MAP_WRITE)) {
    gst_memory_unmap (memory, &in);
    gst_memory_unref (copy);
    return NULL;
  }
  memcpy (out.data, in.data, in.size);
  gst_memory_unmap (copy, &out);
  gst_memory_unmap (memory, &in);
  copy->offset = memory->offset + offset;
  copy->size = size;
  return copy;
}

static void
gst_b200vf_allocator_class_init (GstB200vfAllocatorClass * klass)
{
  GstAllocatorClass *ac = GST_ALLOCATOR_CLASS (klass);
  ac->alloc = GST_DEBUG_FUNCPTR (b200vf_allocator_dummy_alloc);
  ac->free = GST_DEBUG_FUNCPTR (b200vf_allocator_free);
  GST_DEBUG_CATEGORY_INIT (b200vf_memory_debug, "b200vfmemory", 0, "B200 HBM memory");
}

static void
gst_b200vf_allocator_init (GstB200vfAllocator * self)
{
  GstAllocator *alloc = GST_ALLOCATOR_CAST (self);
  alloc->mem_type = GST_B200VF_MEMORY_TYPE;
  alloc->mem_map = b200vf_mem_map;
  alloc->mem_unmap = b200vf_mem_unmap;
  alloc->mem_copy = b200vf_mem_copy;
  /* no mem_share: a sub-memory would need its own transfer state */
  GST_OBJECT_FLAG_SET (self, GST_ALLOCATOR_FLAG_CUSTOM_ALLOC);
}

GstAllocator *
gst_b200vf_allocator_new (b200vf_ctx * ctx)
{
  GstB200vfAllocator *self;
  g_return_val_if_fail (ctx != NULL, NULL);
  self = g_object_new (gst_b200vf_allocator_get_type (), NULL);
  self->ctx = ctx;
  gst_object_ref_sink (self);
  return GST_ALLOCATOR_CAST (self);
}

GstMemory *
gst_b200vf_allocator_wrap (GstAllocator * allocator, b200vf_memory * vf)
{
  GstB200vfMemory *m;
  g_return_val_if_fail (allocator != NULL && vf != NULL, NULL);
  m = g_new0 (GstB200vfMemory, 1);
  m->vf = vf;
  gst_memory_init (GST_MEMORY_CAST (m), GST_MEMORY_FLAG_NO_SHARE, allocator, NULL, b200vf_memory_size (vf), 255, 0,
      b200vf_memory_size (vf));
  return GST_MEMORY_CAST (m);
}

GstMemory *
gst_b200vf_allocator_alloc (GstAllocator * allocator, gsize size)
{
  GstB200vfAllocator *self = (GstB200vfAllocator *) allocator;
  b200vf_memory *vf = NULL;
  g_return_val_if_fail (allocator != NULL, NULL);
  if (b200vf_memory_new (self->ctx, size, &vf) != B200VF_OK) {
    GST_ERROR_OBJECT (self, "allocation of %" G_GSIZE_FORMAT " bytes failed: %s", size, b200vf_last_error ());
    return NULL;
  }
  return gst_b200vf_allocator_wrap (allocator, vf);
}

gboolean
gst_is_b200vf_memory (GstMemory * mem)
{
  return mem != NULL && mem->allocator != NULL && mem->allocator->mem_type != NULL &&
      !strcmp (mem->allocator->mem_type, GST_B200VF_MEMORY_TYPE);
}

b200vf_memory *
gst_b200vf_memory_peek (GstMemory * mem)
{
  return gst_is_b200vf_memory (mem) ? ((GstB200vfMemory *) mem)->vf : NULL;
}

b200vf_memory *
gst_b200vf_buffer_peek (GstBuffer * buffer)
{
  if (gst_buffer_n_memory (buffer) != 1) return NULL;
  return gst_b200vf_memory_peek (gst_buffer_peek_memory (buffer, 0));
}

/* ---------------------------------------------------------------- buffer pool */
typedef struct
{
  GstBufferPool parent;
  b200vf_ctx *ctx;
  GstAllocator *allocator;
  b200vf_pool *slab;            /* one HBM slab for max_buffers buffers; NULL: unbounded pool, one allocation per buffer */
  GstVideoInfo info;
  gboolean have_info;
  gsize size;
} GstB200vfBufferPool;

typedef struct
{
  GstBufferPoolClass parent_class;
} GstB200vfBufferPoolClass;

GType gst_b200vf_buffer_pool_get_type (void);
G_DEFINE_TYPE (GstB200vfBufferPool, gst_b200vf_buffer_pool, GST_TYPE_BUFFER_POOL);

static const gchar **
b200vf_pool_get_options (GstBufferPool * pool)
{
  static const gchar *options[] = { GST_BUFFER_POOL_OPTION_VIDEO_META, NULL };
  return options;
}

static gboolean
b200vf_pool_set_config (GstBufferPool * pool, GstStructure * config)
{
  GstB200vfBufferPool *self = (GstB200vfBufferPool *) pool;
  GstCaps *caps = NULL;
  guint size = 0, min_buffers = 0, max_buffers = 0;
  if (!gst_buffer_pool_config_get_params (config, &caps, &size, &min_buffers, &max_buffers)) {
    GST_WARNING_OBJECT (self, "invalid config");
    return FALSE;
  }
  self->have_info = caps != NULL && gst_video_info_from_caps (&self->info, caps);
  /* video/x-bayer (and anything gst_video_info_from_caps does not know) keeps the size the element asked for */
  self->size = self->have_info ? MAX ((gsize) size, GST_VIDEO_INFO_SIZE (&self->info)) : size;
  if (self->size == 0) {
    GST_WARNING_OBJECT (self, "no buffer size in the config");
    return FALSE;
  }
  if (self->slab) {
    b200vf_pool_destroy (self->slab);
    self->slab = NULL;
  }
  if (max_buffers > 0 && b200vf_pool_create (self->ctx, self->size, (int) max_buffers, &self->slab) != B200VF_OK) {
    GST_WARNING_OBJECT (self, "no HBM slab for %u buffers of %" G_GSIZE_FORMAT " bytes: %s", max_buffers, self->size, b200vf_last_error ());
    self->slab = NULL;          /* fall back to one allocation per buffer */
  }
  gst_buffer_pool_config_set_params (config, caps, (guint) self->size, min_buffers, max_buffers);
  return GST_BUFFER_POOL_CLASS (gst_b200vf_buffer_pool_parent_class)->set_config (pool, config);
}

static GstFlowReturn
b200vf_pool_alloc_buffer (GstBufferPool * pool, GstBuffer ** buffer, GstBufferPoolAcquireParams * params)
{
  GstB200vfBufferPool *self = (GstB200vfBufferPool *) pool;
  GstMemory *mem = NULL;
  GstBuffer *buf;
  if (self->slab) {
    b200vf_memory *vf = NULL;
    if (b200vf_pool_acquire_memory (self->slab, &vf) == B200VF_OK) mem = gst_b200vf_allocator_wrap (self->allocator, vf);
  }
  if (!mem) mem = gst_b200vf_allocator_alloc (self->allocator, self->size);
  if (!mem) {
    GST_ERROR_OBJECT (self, "cannot allocate %" G_GSIZE_FORMAT " bytes of HBM", self->size);
    return GST_FLOW_ERROR;
  }
  buf = gst_buffer_new ();
  gst_buffer_append_memory (buf, mem);
  if (self->have_info)          /* default strides / offsets: the layout the kernels take */
    gst_buffer_add_video_meta_full (buf, GST_VIDEO_FRAME_FLAG_NONE, GST_VIDEO_INFO_FORMAT (&self->info),
        GST_VIDEO_INFO_WIDTH (&self->info), GST_VIDEO_INFO_HEIGHT (&self->info), GST_VIDEO_INFO_N_PLANES (&self->info),
        self->info.offset, self->info.stride);
  *buffer = buf;
  return GST_FLOW_OK;
}

static void
b200vf_pool_finalize (GObject * object)
{
  GstB200vfBufferPool *self = (GstB200vfBufferPool *) object;
  if (self->slab) b200vf_pool_destroy (self->slab);
  if (self->allocator) gst_object_unref (self->allocator);
  G_OBJECT_CLASS (gst_b200vf_buffer_pool_parent_class)->finalize (object);
}

static void
gst_b200vf_buffer_pool_class_init (GstB200vfBufferPoolClass * klass)
{
  GObjectClass *oc = G_OBJECT_CLASS (klass);
  GstBufferPoolClass *pc = GST_BUFFER_POOL_CLASS (klass);
  oc->finalize = b200vf_pool_finalize;
  pc->get_options = b200vf_pool_get_options;
  pc->set_config = b200vf_pool_set_config;
  pc->alloc_buffer = b200vf_pool_alloc_buffer;
}

static void
gst_b200vf_buffer_pool_init (GstB200vfBufferPool * self)
{
  self->slab = NULL;
  self->have_info = FALSE;
}

GstBufferPool *
gst_b200vf_buffer_pool_new (b200vf_ctx * ctx)
{
  GstB200vfBufferPool *self;
  g_return_val_if_fail (ctx != NULL, NULL);
  self = g_object_new (gst_b200vf_buffer_pool_get_type (), NULL);
  self->ctx = ctx;
  self->allocator = gst_b200vf_allocator_new (ctx);
  gst_object_ref_sink (self);
  return GST_BUFFER_POOL_CAST (self);
}
