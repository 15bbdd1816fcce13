/* stub: the part of <gst/gst.h> the shells use (see ../README.md) */
#ifndef STUB_GST_H
#define STUB_GST_H
#include <glib-object.h>
typedef guint64 GstClockTime;
#define GST_CLOCK_TIME_NONE ((GstClockTime) -1)
#define GST_CLOCK_TIME_IS_VALID(time) (((GstClockTime) (time)) != GST_CLOCK_TIME_NONE)
#define GST_ROUND_UP_4(num) (((num) + 3) & ~3)
#define GST_VERSION_MAJOR 1
#define GST_VERSION_MINOR 19

typedef struct _GstMiniObject { GType type; gint refcount; gint lockstate; guint flags; } GstMiniObject;
typedef struct _GstObject { GObject object; gint lock; gchar *name; gpointer parent; guint32 flags; } GstObject;
typedef struct _GstObjectClass { GObjectClass parent_class; } GstObjectClass;
#define GST_OBJECT(obj) ((GstObject *) (obj))
#define GST_OBJECT_FLAGS(obj) (GST_OBJECT (obj)->flags)
#define GST_OBJECT_FLAG_SET(obj, flag) (GST_OBJECT_FLAGS (obj) |= (flag))
void gst_stub_object_lock (gpointer obj);
void gst_stub_object_unlock (gpointer obj);
#define GST_OBJECT_LOCK(obj) gst_stub_object_lock (obj)
#define GST_OBJECT_UNLOCK(obj) gst_stub_object_unlock (obj)
gpointer gst_object_ref_sink (gpointer object);
void gst_object_unref (gpointer object);
gboolean gst_object_sync_values (GstObject * object, GstClockTime timestamp);
#define GST_PARAM_CONTROLLABLE (1 << 9)

/* debug */
typedef struct _GstDebugCategory GstDebugCategory;
#define GST_DEBUG_CATEGORY_STATIC(cat) static GstDebugCategory *cat = NULL
GstDebugCategory *gst_stub_debug_category_new (const gchar * name, guint color, const gchar * description);
#define GST_DEBUG_CATEGORY_INIT(cat, name, color, description) (cat) = gst_stub_debug_category_new ((name), (color), (description))
void gst_stub_log (GstDebugCategory * cat, gpointer object, const gchar * format, ...) __attribute__ ((format (printf, 3, 4)));
#define GST_ERROR(...) gst_stub_log (GST_CAT_DEFAULT, NULL, __VA_ARGS__)
#define GST_WARNING(...) gst_stub_log (GST_CAT_DEFAULT, NULL, __VA_ARGS__)
#define GST_ERROR_OBJECT(obj, ...) gst_stub_log (GST_CAT_DEFAULT, (obj), __VA_ARGS__)
#define GST_WARNING_OBJECT(obj, ...) gst_stub_log (GST_CAT_DEFAULT, (obj), __VA_ARGS__)
#define GST_DEBUG_FUNCPTR(ptr) (ptr)
gchar *gst_stub_error_printf (const gchar * format, ...);
void gst_stub_element_message (gpointer el, const gchar * domain, const gchar * code, gchar * text, gchar * debug);
#define GST_ELEMENT_ERROR(el, domain, code, text, debug) \
  gst_stub_element_message ((el), #domain, #code, gst_stub_error_printf text, gst_stub_error_printf debug)

/* caps / structure */
typedef struct _GstCaps GstCaps;
typedef struct _GstStructure GstStructure;
typedef struct _GstCapsFeatures GstCapsFeatures;
typedef enum { GST_CAPS_INTERSECT_ZIG_ZAG = 0, GST_CAPS_INTERSECT_FIRST = 1 } GstCapsIntersectMode;
GstCaps *gst_caps_from_string (const gchar * string);
GstCaps *gst_caps_copy (const GstCaps * caps);
void gst_caps_unref (GstCaps * caps);
guint gst_caps_get_size (const GstCaps * caps);
GstStructure *gst_caps_get_structure (const GstCaps * caps, guint index);
GstCaps *gst_caps_intersect_full (GstCaps * caps1, GstCaps * caps2, GstCapsIntersectMode mode);
void gst_caps_append (GstCaps * caps1, GstCaps * caps2);
GstCapsFeatures *gst_caps_get_features (const GstCaps * caps, guint index);
void gst_caps_set_features_simple (GstCaps * caps, GstCapsFeatures * features);
GstCapsFeatures *gst_caps_features_new (const gchar * feature1, ...);
gboolean gst_caps_features_contains (const GstCapsFeatures * features, const gchar * feature);
void gst_structure_set_name (GstStructure * structure, const gchar * name);
gboolean gst_structure_has_name (const GstStructure * structure, const gchar * name);
void gst_structure_remove_field (GstStructure * structure, const gchar * fieldname);
void gst_structure_remove_fields (GstStructure * structure, const gchar * fieldname, ...);
gboolean gst_structure_get_int (const GstStructure * structure, const gchar * fieldname, gint * value);
const gchar *gst_structure_get_string (const GstStructure * structure, const gchar * fieldname);

/* memory / allocator */
typedef struct _GstAllocator GstAllocator;
typedef struct _GstAllocatorClass GstAllocatorClass;
typedef struct _GstMemory GstMemory;
typedef struct _GstAllocationParams { guint flags; gsize align; gsize prefix; gsize padding; } GstAllocationParams;
typedef enum { GST_MAP_READ = 1 << 0, GST_MAP_WRITE = 1 << 1, GST_MAP_FLAG_LAST = 1 << 16 } GstMapFlags;
#define GST_MAP_READWRITE ((GstMapFlags) (GST_MAP_READ | GST_MAP_WRITE))
typedef enum { GST_MEMORY_FLAG_READONLY = 1 << 1, GST_MEMORY_FLAG_NO_SHARE = 1 << 4 } GstMemoryFlags;
struct _GstMemory { GstMiniObject mini_object; GstAllocator *allocator; GstMemory *parent; gsize maxsize; gsize align; gsize offset; gsize size; };
typedef struct _GstMapInfo { GstMemory *memory; GstMapFlags flags; guint8 *data; gsize size; gsize maxsize; gpointer user_data[4]; } GstMapInfo;
typedef gpointer (*GstMemoryMapFunction) (GstMemory * mem, gsize maxsize, GstMapFlags flags);
typedef void (*GstMemoryUnmapFunction) (GstMemory * mem);
typedef GstMemory *(*GstMemoryCopyFunction) (GstMemory * mem, gssize offset, gssize size);
typedef GstMemory *(*GstMemoryShareFunction) (GstMemory * mem, gssize offset, gssize size);
struct _GstAllocator {
  GstObject object; const gchar *mem_type; GstMemoryMapFunction mem_map; GstMemoryUnmapFunction mem_unmap;
  GstMemoryCopyFunction mem_copy; GstMemoryShareFunction mem_share;
};
struct _GstAllocatorClass {
  GstObjectClass object_class;
  GstMemory *(*alloc) (GstAllocator * allocator, gsize size, GstAllocationParams * params);
  void (*free) (GstAllocator * allocator, GstMemory * memory);
};
typedef enum { GST_ALLOCATOR_FLAG_CUSTOM_ALLOC = 1 << 4 } GstAllocatorFlags;
GType gst_allocator_get_type (void);
#define GST_TYPE_ALLOCATOR (gst_allocator_get_type ())
#define GST_ALLOCATOR_CAST(obj) ((GstAllocator *) (obj))
#define GST_ALLOCATOR_CLASS(klass) ((GstAllocatorClass *) (klass))
#define GST_MEMORY_CAST(mem) ((GstMemory *) (mem))
void gst_memory_init (GstMemory * mem, GstMemoryFlags flags, GstAllocator * allocator, GstMemory * parent, gsize maxsize,
    gsize align, gsize offset, gsize size);
gboolean gst_memory_map (GstMemory * mem, GstMapInfo * info, GstMapFlags flags);
void gst_memory_unmap (GstMemory * mem, GstMapInfo * info);
void gst_memory_unref (GstMemory * memory);

/* buffer / pool */
typedef struct _GstBufferPool GstBufferPool;
typedef struct _GstBufferPoolClass GstBufferPoolClass;
typedef struct _GstBuffer { GstMiniObject mini_object; GstBufferPool *pool; GstClockTime pts; GstClockTime dts; GstClockTime duration; guint64 offset; guint64 offset_end; } GstBuffer;
#define GST_BUFFER_PTS(buf) (((GstBuffer *) (buf))->pts)
#define GST_BUFFER_TIMESTAMP(buf) GST_BUFFER_PTS (buf)
#define GST_BUFFER_DURATION(buf) (((GstBuffer *) (buf))->duration)
GstBuffer *gst_buffer_new (void);
void gst_buffer_append_memory (GstBuffer * buffer, GstMemory * mem);
guint gst_buffer_n_memory (GstBuffer * buffer);
GstMemory *gst_buffer_peek_memory (GstBuffer * buffer, guint idx);
gboolean gst_buffer_map (GstBuffer * buffer, GstMapInfo * info, GstMapFlags flags);
void gst_buffer_unmap (GstBuffer * buffer, GstMapInfo * info);
typedef enum { GST_FLOW_OK = 0, GST_FLOW_NOT_NEGOTIATED = -4, GST_FLOW_ERROR = -5 } GstFlowReturn;
typedef struct _GstBufferPoolAcquireParams { gint format; gint64 start; gint64 stop; guint flags; } GstBufferPoolAcquireParams;
struct _GstBufferPool { GstObject object; gint flushing; gpointer priv; };
struct _GstBufferPoolClass {
  GstObjectClass object_class;
  const gchar **(*get_options) (GstBufferPool * pool);
  gboolean (*set_config) (GstBufferPool * pool, GstStructure * config);
  gboolean (*start) (GstBufferPool * pool);
  gboolean (*stop) (GstBufferPool * pool);
  GstFlowReturn (*acquire_buffer) (GstBufferPool * pool, GstBuffer ** buffer, GstBufferPoolAcquireParams * params);
  GstFlowReturn (*alloc_buffer) (GstBufferPool * pool, GstBuffer ** buffer, GstBufferPoolAcquireParams * params);
  void (*reset_buffer) (GstBufferPool * pool, GstBuffer * buffer);
  void (*release_buffer) (GstBufferPool * pool, GstBuffer * buffer);
  void (*free_buffer) (GstBufferPool * pool, GstBuffer * buffer);
};
GType gst_buffer_pool_get_type (void);
#define GST_TYPE_BUFFER_POOL (gst_buffer_pool_get_type ())
#define GST_BUFFER_POOL_CAST(obj) ((GstBufferPool *) (obj))
#define GST_BUFFER_POOL_CLASS(klass) ((GstBufferPoolClass *) (klass))
GstStructure *gst_buffer_pool_get_config (GstBufferPool * pool);
gboolean gst_buffer_pool_set_config (GstBufferPool * pool, GstStructure * config);
void gst_buffer_pool_config_set_params (GstStructure * config, GstCaps * caps, guint size, guint min_buffers, guint max_buffers);
gboolean gst_buffer_pool_config_get_params (GstStructure * config, GstCaps ** caps, guint * size, guint * min_buffers, guint * max_buffers);
void gst_buffer_pool_config_add_option (GstStructure * config, const gchar * option);

/* query */
typedef struct _GstQuery GstQuery;
void gst_query_parse_allocation (GstQuery * query, GstCaps ** caps, gboolean * need_pool);
void gst_query_add_allocation_pool (GstQuery * query, GstBufferPool * pool, guint size, guint min_buffers, guint max_buffers);
guint gst_query_get_n_allocation_pools (GstQuery * query);
void gst_query_parse_nth_allocation_pool (GstQuery * query, guint index, GstBufferPool ** pool, guint * size, guint * min_buffers, guint * max_buffers);
void gst_query_set_nth_allocation_pool (GstQuery * query, guint index, GstBufferPool * pool, guint size, guint min_buffers, guint max_buffers);
void gst_query_add_allocation_meta (GstQuery * query, GType api, const GstStructure * params);

/* element / pad / plugin */
typedef struct _GstElement { GstObject object; gpointer priv[16]; } GstElement;
typedef struct _GstElementClass { GstObjectClass parent_class; gpointer priv[32]; } GstElementClass;
#define GST_ELEMENT_CLASS(klass) ((GstElementClass *) (klass))
typedef struct _GstPad GstPad;
typedef struct _GstPadTemplate GstPadTemplate;
typedef struct _GstEvent GstEvent;
typedef struct _GstPlugin GstPlugin;
typedef enum { GST_PAD_UNKNOWN, GST_PAD_SRC, GST_PAD_SINK } GstPadDirection;
typedef enum { GST_PAD_ALWAYS, GST_PAD_SOMETIMES, GST_PAD_REQUEST } GstPadPresence;
typedef enum { GST_RANK_NONE = 0 } GstRank;
typedef enum { GST_FORMAT_UNDEFINED = 0, GST_FORMAT_TIME = 3 } GstFormat;
typedef struct _GstSegment { guint flags; gdouble rate; gdouble applied_rate; GstFormat format; guint64 base, offset, start, stop, time, position, duration; } GstSegment;
guint64 gst_segment_to_stream_time (const GstSegment * segment, GstFormat format, guint64 position);
guint64 gst_segment_to_running_time (const GstSegment * segment, GstFormat format, guint64 position);
GstPadTemplate *gst_pad_template_new (const gchar * name_template, GstPadDirection direction, GstPadPresence presence, GstCaps * caps);
void gst_element_class_add_pad_template (GstElementClass * klass, GstPadTemplate * templ);
void gst_element_class_set_static_metadata (GstElementClass * klass, const gchar * longname, const gchar * classification,
    const gchar * description, const gchar * author);
gboolean gst_element_register (GstPlugin * plugin, const gchar * name, guint rank, GType type);
gboolean gst_pad_push_event (GstPad * pad, GstEvent * event);
typedef struct _GstMessage GstMessage;
GstStructure *gst_structure_new (const gchar * name, const gchar * firstfield, ...);
GstMessage *gst_message_new_element (GstObject * src, GstStructure * structure);
gboolean gst_element_post_message (GstElement * element, GstMessage * message);
#define GST_ELEMENT_CAST(obj) ((GstElement *) (obj))
#define GST_OBJECT_CAST(obj) ((GstObject *) (obj))
typedef gboolean (*GstPluginInitFunc) (GstPlugin * plugin);
typedef struct _GstPluginDesc {
  gint major_version, minor_version; const gchar *name, *description; GstPluginInitFunc plugin_init;
  const gchar *version, *license, *source, *package, *origin, *release_datetime;
} GstPluginDesc;
#define GST_PLUGIN_DEFINE(major, minor, name, description, init, version, license, package, origin) \
  const GstPluginDesc gst_plugin_##name##_desc = { major, minor, #name, description, init, version, license, PACKAGE, package, origin, NULL };
#endif
