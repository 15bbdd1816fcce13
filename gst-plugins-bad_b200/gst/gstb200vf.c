/* gstb200vf.c - GStreamer element shells over libb200vf.so (C/GLib, no per-pixel work).
 *
 * One source, compiled once per plugin with
 * -DB200VF_PLUGIN=<bayer|gaudieffects|coloreffects|geometrictransform|videofiltersbad|smooth|videosignal>
 * (gst/meson.build), yields libgstbayer.so, libgstgaudieffects.so, libgstcoloreffects.so,
 * libgstgeometrictransform.so and libgstvideofiltersbad.so (zebrastripe, videodiff, scenechange) that REPLACE the stock plugins: same plugin names, factory names, GType names and
 * parents, klass/description strings, GObject properties (names, ranges, defaults, GST_PARAM_CONTROLLABLE) and pad
 * templates - all taken from the introspection table of the library (b200vf_factory_*), which is generated from and
 * tested against the reference's own API dump (docs/plugins/gst_plugins_cache.json). The data vfuncs
 * (GstBaseTransform::transform, gst/bayer/gstbayer2rgb.c:456-487; GstVideoFilter::transform_frame[_ip],
 * e.g. gst/gaudieffects/gstburn.c:214-250) take one of two paths. Buffers of the HBM pool (GstB200vfMemory, negotiated
 * through the caps feature memory:B200VFMemory and propose / decide_allocation - none of the reference elements
 * overrides those, SURVEY 8b: the default sysmem pool is what this replaces) go to b200vf_element_transform: the frame
 * stays in HBM from element to element and per-pixel elements fuse. System-memory buffers are mapped and handed to
 * b200vf_element_transform_host (page-locked in place on first sight), with the frame's own plane strides / offsets
 * when a GstVideoMeta moved them (b200vf_element_transform_host_layout). start() fails with a GST_ELEMENT_ERROR when there is no sm_100 device (no CPU path), like
 * sys/nvcodec registers nothing without a driver (sys/nvcodec/plugin.c:72-103).
 *
 * There is no GLib / GStreamer in this repository's CI image: the file is compile-checked there against
 * the declaration-only headers of tests/stubs/ (tests/test_shells_cpu.py: every plugin variant, -Wall -Werror) and built
 * for real by gst/meson.build where `pkg-config gstreamer-video-1.0` exists. The mapping from element state to C-ABI
 * calls is the one host/elements.cpp exercises on the GPU in tests/test_elements_gpu.py / test_memory_gpu.py.
 */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#include <string.h>
#include <gst/gst.h>
#include <gst/base/gstbasetransform.h>
#include <gst/video/video.h>
#include <gst/video/gstvideofilter.h>
#include "b200vf.h"
#include "gstb200vfmemory.h"

#ifndef B200VF_PLUGIN
#error "compile with -DB200VF_PLUGIN=bayer|gaudieffects|coloreffects|geometrictransform|videofiltersbad"
#endif
#define STR_(x) #x
#define STR(x) STR_(x)
#ifndef PACKAGE
#define PACKAGE "gst-plugins-bad"
#endif
#ifndef VERSION
#define VERSION "1.19.2"
#endif

GST_DEBUG_CATEGORY_STATIC (b200vf_debug);
#define GST_CAT_DEFAULT b200vf_debug

typedef struct
{
  GstVideoFilter parent;        /* GstBaseTransform is its first member: bayer elements use only that part */
  b200vf_ctx *ctx;
  b200vf_element *el;
  gint device;
  guint key_unit_count;         /* scenechange: running count of the force-key-unit events (gstscenechange.c:255) */
  gboolean hbm_out;             /* decide_allocation settled on the HBM pool for the src side */
} GstB200vf;

typedef struct
{
  GstVideoFilterClass parent_class;
  const gchar *factory;         /* set from class_data */
} GstB200vfClass;

#define B200VF(obj) ((GstB200vf *) (obj))
#define B200VF_GET_CLASS(obj) ((GstB200vfClass *) G_OBJECT_GET_CLASS (obj))

enum { PROP_0, PROP_FIRST };    /* property ids: PROP_FIRST + index in the library's table */

static gboolean
is_bayer_plugin_factory (const gchar * f)
{
  return !strcmp (f, "bayer2rgb") || !strcmp (f, "rgb2bayer");
}

/* ---- properties: straight through to the element mirror (which validates ranges like GParamSpec does) ---- */
/* perspective's only own property is `matrix`, a GValueArray of 9 doubles (gstperspective.c:97-176, 222-233); the mirror
 * keeps the nine coefficients as matrix-0 .. matrix-8 */
static gboolean
is_matrix_property (GParamSpec * pspec)
{
  return !strcmp (pspec->name, "matrix");
}

static void
b200vf_set_property (GObject * object, guint id, const GValue * value, GParamSpec * pspec)
{
  GstB200vf *self = B200VF (object);
  gdouble v = 0;
  if (is_matrix_property (pspec)) {
    GValueArray *va = g_value_get_boxed (value);
    if (!va || va->n_values != 9) {
      GST_WARNING_OBJECT (self, "matrix: invalid number of elements: %u", va ? va->n_values : 0);     /* set_matrix_from_array, :114-135 */
      return;
    }
    GST_OBJECT_LOCK (self);
    for (guint i = 0; i < 9; i++) {
      gchar name[16];
      g_snprintf (name, sizeof name, "matrix-%u", i);
      b200vf_element_set_property (self->el, name, g_value_get_double (g_value_array_get_nth (va, i)));
    }
    GST_OBJECT_UNLOCK (self);
    return;
  }
  if (G_VALUE_HOLDS_UINT (value)) v = g_value_get_uint (value);
  else if (G_VALUE_HOLDS_UINT64 (value)) v = (gdouble) g_value_get_uint64 (value);
  else if (G_VALUE_HOLDS_INT (value)) v = g_value_get_int (value);
  else if (G_VALUE_HOLDS_BOOLEAN (value)) v = g_value_get_boolean (value);
  else if (G_VALUE_HOLDS_DOUBLE (value)) v = g_value_get_double (value);
  else if (G_VALUE_HOLDS_ENUM (value)) v = g_value_get_enum (value);
  GST_OBJECT_LOCK (self);
  if (b200vf_element_set_property (self->el, pspec->name, v) != B200VF_OK)
    GST_WARNING_OBJECT (self, "%s", b200vf_last_error ());
  GST_OBJECT_UNLOCK (self);
}

static void
b200vf_get_property (GObject * object, guint id, GValue * value, GParamSpec * pspec)
{
  GstB200vf *self = B200VF (object);
  gdouble v = 0;
  if (is_matrix_property (pspec)) {         /* get_array_from_matrix, :96-112 */
    GValueArray *va = g_value_array_new (1);
    for (guint i = 0; i < 9; i++) {
      GValue d = G_VALUE_INIT;
      gchar name[16];
      g_snprintf (name, sizeof name, "matrix-%u", i);
      b200vf_element_get_property (self->el, name, &v);
      g_value_init (&d, G_TYPE_DOUBLE);
      g_value_set_double (&d, v);
      g_value_array_append (va, &d);
      g_value_unset (&d);
    }
    g_value_take_boxed (value, va);
    return;
  }
  b200vf_element_get_property (self->el, pspec->name, &v);
  if (G_VALUE_HOLDS_UINT (value)) g_value_set_uint (value, (guint) v);
  else if (G_VALUE_HOLDS_UINT64 (value)) g_value_set_uint64 (value, (guint64) v);
  else if (G_VALUE_HOLDS_INT (value)) g_value_set_int (value, (gint) v);
  else if (G_VALUE_HOLDS_BOOLEAN (value)) g_value_set_boolean (value, v != 0);
  else if (G_VALUE_HOLDS_DOUBLE (value)) g_value_set_double (value, v);
  else if (G_VALUE_HOLDS_ENUM (value)) g_value_set_enum (value, (gint) v);
}

static GType
enum_type_for (const gchar * type_name, const b200vf_property_info * p)
{
  /* GstColorEffectsPreset, GstMirrorMode, GstGeometricTransformOffEdgesPixelsMethod (reference type names) */
  const gchar *name = !strcmp (p->name, "preset") ? "GstColorEffectsPreset" :
      !strcmp (p->name, "mode") ? "GstMirrorMode" : "GstGeometricTransformOffEdgesPixelsMethod";
  GType t = g_type_from_name (name);
  if (!t) {
    GEnumValue *vals = g_new0 (GEnumValue, p->n_nicks + 1);
    for (gint i = 0; i < p->n_nicks; i++) {
      vals[i].value = i;
      vals[i].value_name = p->nicks[i];
      vals[i].value_nick = p->nicks[i];
    }
    t = g_enum_register_static (name, vals);
  }
  return t;
}

/* ---- negotiation ---- */
static gboolean
b200vf_configure (GstB200vf * self, const gchar * in_fmt, const gchar * out_fmt, gint w, gint h)
{
  if (b200vf_element_set_caps (self->el, in_fmt, out_fmt, w, h) != B200VF_OK) {
    GST_WARNING_OBJECT (self, "%s", b200vf_last_error ());
    return FALSE;
  }
  return TRUE;
}

/* GstVideoFilter::set_info (e.g. gstgaussblur.c:162-178, gstgeometrictransform.c:130-165) */
static gboolean
b200vf_set_info (GstVideoFilter * vf, GstCaps * incaps, GstVideoInfo * in_info, GstCaps * outcaps, GstVideoInfo * out_info)
{
  const gchar *f = gst_video_format_to_string (GST_VIDEO_INFO_FORMAT (in_info));
  return b200vf_configure (B200VF (vf), f, f, GST_VIDEO_INFO_WIDTH (in_info), GST_VIDEO_INFO_HEIGHT (in_info));
}

/* bayer2rgb / rgb2bayer: set_caps, transform_caps, get_unit_size (gstbayer2rgb.c:237-352) */
static gboolean
b200vf_bayer_set_caps (GstBaseTransform * base, GstCaps * incaps, GstCaps * outcaps)
{
  GstStructure *si = gst_caps_get_structure (incaps, 0), *so = gst_caps_get_structure (outcaps, 0);
  gint w = 0, h = 0;
  gst_structure_get_int (si, "width", &w);
  gst_structure_get_int (si, "height", &h);
  return b200vf_configure (B200VF (base), gst_structure_get_string (si, "format"), gst_structure_get_string (so, "format"), w, h);
}

static GstCaps *
b200vf_bayer_transform_caps (GstBaseTransform * base, GstPadDirection direction, GstCaps * caps, GstCaps * filter)
{
  const gchar *factory = B200VF_GET_CLASS (base)->factory;
  gboolean to_raw = (!strcmp (factory, "bayer2rgb")) == (direction == GST_PAD_SINK);
  GstCaps *res = gst_caps_copy (caps);
  for (guint i = 0; i < gst_caps_get_size (res); i++) {
    GstStructure *s = gst_caps_get_structure (res, i);
    gst_structure_set_name (s, to_raw ? "video/x-raw" : "video/x-bayer");
    if (to_raw) gst_structure_remove_field (s, "format");
    else gst_structure_remove_fields (s, "format", "colorimetry", "chroma-site", NULL);
  }
  if (filter) {
    GstCaps *tmp = res;
    res = gst_caps_intersect_full (filter, tmp, GST_CAPS_INTERSECT_FIRST);
    gst_caps_unref (tmp);
  }
  return res;
}

static gboolean
b200vf_bayer_get_unit_size (GstBaseTransform * base, GstCaps * caps, gsize * size)
{
  GstStructure *s = gst_caps_get_structure (caps, 0);
  gint w, h;
  if (!gst_structure_get_int (s, "width", &w) || !gst_structure_get_int (s, "height", &h)) {
    GST_ELEMENT_ERROR (base, CORE, NEGOTIATION, (NULL), ("Incomplete caps, some required field missing"));
    return FALSE;
  }
  *size = gst_structure_has_name (s, "video/x-raw") ? (gsize) w * h * 4 : (gsize) GST_ROUND_UP_4 (w) * h;
  return TRUE;
}

/* ---- lifecycle ---- */
static gboolean
b200vf_start (GstBaseTransform * base)
{
  GstB200vf *self = B200VF (base);
  if (!self->ctx && b200vf_ctx_create (self->device, &self->ctx) != B200VF_OK) {
    GST_ELEMENT_ERROR (self, LIBRARY, INIT, ("No B200 (sm_100) device: this element has no CPU path"), ("%s", b200vf_last_error ()));
    return FALSE;
  }
  /* re-create the mirror with a device context, carrying the property values over */
  b200vf_element *old = self->el;
  if (b200vf_element_factory_make (self->ctx, B200VF_GET_CLASS (self)->factory, &self->el) != B200VF_OK) {
    self->el = old;
    return FALSE;
  }
  b200vf_factory_info fi;
  b200vf_factory_find (B200VF_GET_CLASS (self)->factory, &fi);
  for (gint i = 0; i < fi.n_properties; i++) {
    b200vf_property_info p;
    gdouble v;
    b200vf_factory_property (fi.factory, i, &p);
    if (b200vf_element_get_property (old, p.name, &v) == B200VF_OK) b200vf_element_set_property (self->el, p.name, v);
  }
  b200vf_element_destroy (old);
  /* sysmem GstBuffers are pageable and recur (buffer pools): page-lock each range on first sight */
  b200vf_element_set_host_mode (self->el, 1);
  return TRUE;
}

static gboolean
b200vf_stop (GstBaseTransform * base)
{
  return TRUE;                  /* the index table / staging buffers live with the element mirror */
}

static void
b200vf_before_transform (GstBaseTransform * base, GstBuffer * buf)
{
  /* GstController: update the properties (gstburn.c:230-240, gstgeometrictransform.c:209-224) */
  GstClockTime ts = gst_segment_to_stream_time (&base->segment, GST_FORMAT_TIME, GST_BUFFER_TIMESTAMP (buf));
  if (GST_CLOCK_TIME_IS_VALID (ts)) gst_object_sync_values (GST_OBJECT (base), ts);
}

/* ---- data ---- */
static GstFlowReturn
b200vf_flow (GstB200vf * self, int rc)
{
  if (rc == B200VF_OK) return GST_FLOW_OK;
  if (rc == B200VF_E_NOT_NEGOTIATED) return GST_FLOW_NOT_NEGOTIATED;
  GST_ELEMENT_ERROR (self, LIBRARY, FAILED, ("b200vf: %s", b200vf_status_string (rc)), ("%s", b200vf_last_error ()));
  return GST_FLOW_ERROR;
}

/* the frame's plane strides / offsets against the layout the library assumes; *differs = the frame needs
 * b200vf_element_transform_host_layout */
static void
frame_layout (GstB200vf * self, int side, GstVideoFrame * frame, b200vf_frame_layout * lay, gboolean * differs)
{
  b200vf_frame_layout def;
  guint8 *base = GST_VIDEO_FRAME_PLANE_DATA (frame, 0);
  memset (lay, 0, sizeof *lay);
  *differs = FALSE;
  if (b200vf_element_default_layout (self->el, side, &def) != B200VF_OK) return;
  lay->n_planes = (int) GST_VIDEO_FRAME_N_PLANES (frame);
  if (lay->n_planes != def.n_planes) *differs = TRUE;
  for (gint i = 0; i < lay->n_planes && i < 4; i++) {
    lay->offset[i] = (size_t) ((guint8 *) GST_VIDEO_FRAME_PLANE_DATA (frame, i) - base);
    lay->stride[i] = GST_VIDEO_FRAME_PLANE_STRIDE (frame, i);
    if (i < def.n_planes && (lay->offset[i] != def.offset[i] || lay->stride[i] != def.stride[i])) *differs = TRUE;
  }
}

static void
push_scenechange_event (GstB200vf * self, GstBuffer * buf)
{
  /* scenechange: the detection becomes a downstream force-key-unit event (gstscenechange.c:246-257) */
  int changed = 0;
  if (strcmp (B200VF_GET_CLASS (self)->factory, "scenechange")) return;
  if (b200vf_element_last_events (self->el, &changed, 1) == 1 && changed)
    gst_pad_push_event (GST_BASE_TRANSFORM_SRC_PAD (self),
        gst_video_event_new_downstream_force_key_unit (GST_BUFFER_PTS (buf), GST_CLOCK_TIME_NONE,
            GST_CLOCK_TIME_NONE, FALSE, self->key_unit_count++));
}

/* videoanalyse / simplevideomarkdetect report through element messages (gstvideoanalyse.c:178-204,
 * gstsimplevideomarkdetect.c:352-389); the numbers come back from the library (b200vf_element_last_values) */
static void
post_analysis_message (GstB200vf * self, GstBuffer * buf)
{
  GstBaseTransform *trans = GST_BASE_TRANSFORM (self);
  const gchar *factory = B200VF_GET_CLASS (self)->factory;
  gboolean analyse = !strcmp (factory, "videoanalyse"), detect = !strcmp (factory, "simplevideomarkdetect");
  gdouble v[3] = { 0, 0, 0 }, want_message = 1;
  guint64 duration, timestamp, running_time, stream_time;
  GstStructure *st;
  if (!analyse && !detect) return;
  b200vf_element_get_property (self->el, "message", &want_message);
  if (want_message == 0 || b200vf_element_last_values (self->el, v, 3) < 2) return;
  if (detect && v[0] == 0) return;                /* the detector posts only when a pattern appears, changes or disappears */
  timestamp = GST_BUFFER_TIMESTAMP (buf);
  duration = GST_BUFFER_DURATION (buf);
  running_time = gst_segment_to_running_time (&trans->segment, GST_FORMAT_TIME, timestamp);
  stream_time = gst_segment_to_stream_time (&trans->segment, GST_FORMAT_TIME, timestamp);
  if (analyse)
    st = gst_structure_new ("GstVideoAnalyse", "timestamp", G_TYPE_UINT64, timestamp, "stream-time", G_TYPE_UINT64, stream_time,
        "running-time", G_TYPE_UINT64, running_time, "duration", G_TYPE_UINT64, duration, "luma-average", G_TYPE_DOUBLE, v[0],
        "luma-variance", G_TYPE_DOUBLE, v[1], NULL);
  else
    st = gst_structure_new ("GstSimpleVideoMarkDetect", "have-pattern", G_TYPE_BOOLEAN, v[1] != 0, "timestamp", G_TYPE_UINT64, timestamp,
        "stream-time", G_TYPE_UINT64, stream_time, "running-time", G_TYPE_UINT64, running_time, "duration", G_TYPE_UINT64, duration,
        "data", G_TYPE_UINT64, (guint64) v[2], NULL);
  gst_element_post_message (GST_ELEMENT_CAST (self), gst_message_new_element (GST_OBJECT_CAST (self), st));
}

/* GstVideoFilter elements: GstBaseTransform::transform / transform_ip are overridden (instead of transform_frame[_ip])
 * because GstVideoFilter maps both buffers to system memory before it calls transform_frame - for buffers of the HBM
 * pool that would download every frame. */
static GstFlowReturn
b200vf_vf_transform (GstBaseTransform * base, GstBuffer * inbuf, GstBuffer * outbuf)
{
  GstB200vf *self = B200VF (base);
  GstVideoFilter *vf = GST_VIDEO_FILTER_CAST (base);
  b200vf_memory *min = gst_b200vf_buffer_peek (inbuf), *mout = gst_b200vf_buffer_peek (outbuf);
  GstVideoFrame fin, fout;
  b200vf_frame_layout lin, lout;
  gboolean din, dout;
  int rc;
  if (!vf->negotiated) return GST_FLOW_NOT_NEGOTIATED;
  if (min && mout)              /* HBM in, HBM out: nothing crosses PCIe, per-pixel elements only record themselves */
    return b200vf_flow (self, b200vf_element_transform (self->el, min, mout, 1, NULL));
  if (!gst_video_frame_map (&fin, &vf->in_info, inbuf, GST_MAP_READ)) goto map_failed;
  if (!gst_video_frame_map (&fout, &vf->out_info, outbuf, GST_MAP_WRITE)) {
    gst_video_frame_unmap (&fin);
    goto map_failed;
  }
  frame_layout (self, 0, &fin, &lin, &din);
  frame_layout (self, 1, &fout, &lout, &dout);
  if (din || dout)
    rc = b200vf_element_transform_host_layout (self->el, GST_VIDEO_FRAME_PLANE_DATA (&fin, 0), &lin, GST_VIDEO_FRAME_PLANE_DATA (&fout, 0), &lout);
  else
    rc = b200vf_element_transform_host (self->el, GST_VIDEO_FRAME_PLANE_DATA (&fin, 0), GST_VIDEO_FRAME_PLANE_DATA (&fout, 0), 1);
  gst_video_frame_unmap (&fout);
  gst_video_frame_unmap (&fin);
  return b200vf_flow (self, rc);
map_failed:
  GST_ELEMENT_ERROR (self, CORE, FAILED, (NULL), ("could not map a video frame"));
  return GST_FLOW_ERROR;
}

static GstFlowReturn
b200vf_vf_transform_ip (GstBaseTransform * base, GstBuffer * buf)
{
  GstB200vf *self = B200VF (base);
  GstVideoFilter *vf = GST_VIDEO_FILTER_CAST (base);
  b200vf_memory *m = gst_b200vf_buffer_peek (buf);
  GstVideoFrame frame;
  b200vf_frame_layout lay;
  gboolean differs;
  GstFlowReturn ret;
  int rc;
  if (!vf->negotiated) return GST_FLOW_NOT_NEGOTIATED;
  if (m) {
    ret = b200vf_flow (self, b200vf_element_transform (self->el, m, m, 1, NULL));
  } else {
    guint8 *d;
    if (!gst_video_frame_map (&frame, &vf->in_info, buf, GST_MAP_READWRITE)) {
      GST_ELEMENT_ERROR (self, CORE, FAILED, (NULL), ("could not map a video frame"));
      return GST_FLOW_ERROR;
    }
    frame_layout (self, 0, &frame, &lay, &differs);
    d = GST_VIDEO_FRAME_PLANE_DATA (&frame, 0);
    rc = differs ? b200vf_element_transform_host_layout (self->el, d, &lay, d, &lay) : b200vf_element_transform_host (self->el, d, d, 1);
    gst_video_frame_unmap (&frame);
    ret = b200vf_flow (self, rc);
  }
  if (ret == GST_FLOW_OK) {
    push_scenechange_event (self, buf);
    post_analysis_message (self, buf);
  }
  return ret;
}

static GstFlowReturn
b200vf_bayer_transform (GstBaseTransform * base, GstBuffer * inbuf, GstBuffer * outbuf)
{
  GstMapInfo in, out;
  b200vf_memory *min = gst_b200vf_buffer_peek (inbuf), *mout = gst_b200vf_buffer_peek (outbuf);
  if (min && mout) return b200vf_flow (B200VF (base), b200vf_element_transform (B200VF (base)->el, min, mout, 1, NULL));
  if (!gst_buffer_map (inbuf, &in, GST_MAP_READ)) goto map_failed;
  if (!gst_buffer_map (outbuf, &out, GST_MAP_WRITE)) {
    gst_buffer_unmap (inbuf, &in);
    goto map_failed;
  }
  {
    /* the mosaic has no GstVideoInfo: its pitch is GST_ROUND_UP_4 (width) by the element's own contract (:477); the RGB
     * side takes whatever stride a GstVideoMeta announces */
    GstVideoMeta *meta = gst_buffer_get_video_meta (outbuf);
    b200vf_frame_layout lo;
    int rc;
    gboolean raw_out = !strcmp (B200VF_GET_CLASS (base)->factory, "bayer2rgb");
    if (meta && raw_out && b200vf_element_default_layout (B200VF (base)->el, 1, &lo) == B200VF_OK && meta->n_planes == 1 &&
        (meta->stride[0] != lo.stride[0] || meta->offset[0] != 0)) {
      lo.stride[0] = meta->stride[0];
      lo.offset[0] = meta->offset[0];
      rc = b200vf_element_transform_host_layout (B200VF (base)->el, in.data, NULL, out.data, &lo);
    } else
      rc = b200vf_element_transform_host (B200VF (base)->el, in.data, out.data, 1);
    gst_buffer_unmap (outbuf, &out);
    gst_buffer_unmap (inbuf, &in);
    return b200vf_flow (B200VF (base), rc);
  }
map_failed:
  GST_WARNING_OBJECT (base, "Could not map buffer, skipping");      /* the reference returns OK here (:484-486) */
  return GST_FLOW_OK;
}

/* ---- allocation: the HBM pool replaces the default system-memory pool (sys/nvcodec/gstcudabasetransform.c:436-587) ---- */
static gboolean
caps_have_hbm_feature (GstCaps * caps)
{
  GstCapsFeatures *f = (caps && gst_caps_get_size (caps) > 0) ? gst_caps_get_features (caps, 0) : NULL;
  return f != NULL && gst_caps_features_contains (f, GST_CAPS_FEATURE_MEMORY_B200VF);
}

static gboolean
b200vf_propose_allocation (GstBaseTransform * base, GstQuery * decide_query, GstQuery * query)
{
  GstB200vf *self = B200VF (base);
  GstCaps *caps = NULL;
  gboolean need_pool = FALSE;
  if (decide_query == NULL) return TRUE;      /* passthrough: upstream and downstream talk to each other */
  gst_query_parse_allocation (query, &caps, &need_pool);
  if (caps == NULL) return FALSE;
  if (caps_have_hbm_feature (caps) && self->ctx && need_pool) {
    /* upstream produces into HBM: offer it our pool */
    gsize in_bytes = 0, out_bytes = 0;
    GstBufferPool *pool = gst_b200vf_buffer_pool_new (self->ctx);
    GstStructure *config = gst_buffer_pool_get_config (pool);
    b200vf_element_unit_size (self->el, &in_bytes, &out_bytes);
    gst_buffer_pool_config_set_params (config, caps, (guint) in_bytes, 0, 0);
    gst_buffer_pool_config_add_option (config, GST_BUFFER_POOL_OPTION_VIDEO_META);
    if (!gst_buffer_pool_set_config (pool, config)) {
      gst_object_unref (pool);
      return FALSE;
    }
    gst_query_add_allocation_pool (query, pool, (guint) in_bytes, 0, 0);
    gst_object_unref (pool);
  }
  /* strides other than the default ones are fine with us (transform_host_layout) */
  gst_query_add_allocation_meta (query, GST_VIDEO_META_API_TYPE, NULL);
  return TRUE;
}

static gboolean
b200vf_decide_allocation (GstBaseTransform * base, GstQuery * query)
{
  GstB200vf *self = B200VF (base);
  GstCaps *outcaps = NULL;
  GstBufferPool *pool = NULL;
  guint size = 0, min = 0, max = 0;
  gboolean update_pool = FALSE;
  gst_query_parse_allocation (query, &outcaps, NULL);
  self->hbm_out = outcaps != NULL && caps_have_hbm_feature (outcaps) && self->ctx != NULL;
  if (!self->hbm_out)           /* downstream wants system memory: the base class's default video pool */
    return GST_BASE_TRANSFORM_CLASS (g_type_class_peek_parent (G_OBJECT_GET_CLASS (base)))->decide_allocation (base, query);
  if (gst_query_get_n_allocation_pools (query) > 0) {
    gst_query_parse_nth_allocation_pool (query, 0, &pool, &size, &min, &max);
    update_pool = TRUE;
    if (pool && !g_type_is_a (G_OBJECT_TYPE (pool), g_type_from_name ("GstB200vfBufferPool"))) {
      gst_object_unref (pool);  /* somebody else's pool: ours allocates in HBM */
      pool = NULL;
    }
  }
  {
    gsize in_bytes = 0, out_bytes = 0;
    GstStructure *config;
    b200vf_element_unit_size (self->el, &in_bytes, &out_bytes);
    if (size < out_bytes) size = (guint) out_bytes;
    if (!pool) pool = gst_b200vf_buffer_pool_new (self->ctx);
    config = gst_buffer_pool_get_config (pool);
    gst_buffer_pool_config_set_params (config, outcaps, size, min, max);
    gst_buffer_pool_config_add_option (config, GST_BUFFER_POOL_OPTION_VIDEO_META);
    gst_buffer_pool_set_config (pool, config);
  }
  if (update_pool) gst_query_set_nth_allocation_pool (query, 0, pool, size, min, max);
  else gst_query_add_allocation_pool (query, pool, size, min, max);
  gst_object_unref (pool);
  return TRUE;
}

/* ---- class / instance init, driven by the introspection table ---- */
static void
b200vf_finalize (GObject * object)
{
  GstB200vf *self = B200VF (object);
  if (self->el) b200vf_element_destroy (self->el);
  if (self->ctx) b200vf_ctx_destroy (self->ctx);
  G_OBJECT_CLASS (g_type_class_peek_parent (G_OBJECT_GET_CLASS (object)))->finalize (object);
}

/* template caps: the reference's system-memory caps first (what an unchanged pipeline negotiates), then the same
 * with the memory:B200VFMemory feature (sys/nvcodec/gstcudabasetransform.c:57-106 does the same for CUDAMemory) */
static GstCaps *
with_hbm_feature (GstCaps * caps)
{
  GstCaps *hbm = gst_caps_copy (caps);
  gst_caps_set_features_simple (hbm, gst_caps_features_new (GST_CAPS_FEATURE_MEMORY_B200VF, NULL));
  gst_caps_append (caps, hbm);
  return caps;
}

static GstCaps *
raw_caps_for (const b200vf_factory_info * fi)
{
  GString *s = g_string_new ("{ ");
  for (gint i = 0; i < fi->n_formats; i++) g_string_append_printf (s, "%s%s", i ? ", " : "", b200vf_factory_format (fi->factory, i));
  g_string_append (s, " }");
  gchar *c = g_strdup_printf (GST_VIDEO_CAPS_MAKE ("%s"), s->str);
  GstCaps *caps = gst_caps_from_string (c);
  g_free (c);
  g_string_free (s, TRUE);
  return with_hbm_feature (caps);
}

static void
b200vf_class_init (gpointer g_class, gpointer class_data)
{
  GstB200vfClass *klass = g_class;
  GObjectClass *oc = G_OBJECT_CLASS (g_class);
  GstElementClass *ec = GST_ELEMENT_CLASS (g_class);
  GstBaseTransformClass *bc = GST_BASE_TRANSFORM_CLASS (g_class);
  b200vf_factory_info fi;
  klass->factory = class_data;
  b200vf_factory_find (klass->factory, &fi);

  oc->set_property = b200vf_set_property;
  oc->get_property = b200vf_get_property;
  oc->finalize = b200vf_finalize;
  gst_element_class_set_static_metadata (ec, fi.long_name, fi.klass, fi.description, fi.author);

  for (gint i = 0; i < fi.n_properties; i++) {
    b200vf_property_info p;
    GParamSpec *ps = NULL;
    GParamFlags fl = G_PARAM_READWRITE | G_PARAM_STATIC_STRINGS;
    b200vf_factory_property (fi.factory, i, &p);
    if (p.controllable) fl |= GST_PARAM_CONTROLLABLE;
    if (!strncmp (p.name, "matrix-", 7)) {
      /* perspective: the nine coefficients are ONE GObject property, "matrix", a GValueArray of doubles
       * (gstperspective.c:222-233); installed once, under the id of matrix-0 */
      if (!strcmp (p.name, "matrix-0"))
        g_object_class_install_property (oc, PROP_FIRST + i, g_param_spec_value_array ("matrix", "Matrix",
                "Matrix of dimension 3x3 to use in the 2D transform, passed as an array of 9 elements in row-major order",
                g_param_spec_double ("Element", "Transformation matrix element", "Element of the transformation matrix",
                    -G_MAXDOUBLE, G_MAXDOUBLE, 0.0, G_PARAM_READWRITE | G_PARAM_STATIC_STRINGS), fl));
      continue;
    }
    switch (p.type) {
      case B200VF_PROP_UINT: ps = g_param_spec_uint (p.name, p.name, p.name, (guint) p.min, (guint) p.max, (guint) p.def, fl); break;
      case B200VF_PROP_INT: ps = g_param_spec_int (p.name, p.name, p.name, (gint) p.min, (gint) p.max, (gint) p.def, fl); break;
      case B200VF_PROP_BOOL: ps = g_param_spec_boolean (p.name, p.name, p.name, p.def != 0, fl); break;
      case B200VF_PROP_DOUBLE: ps = g_param_spec_double (p.name, p.name, p.name, p.min, p.max, p.def, fl); break;
      case B200VF_PROP_UINT64: ps = g_param_spec_uint64 (p.name, p.name, p.name, 0, G_MAXUINT64, (guint64) p.def, fl); break;
      case B200VF_PROP_ENUM: ps = g_param_spec_enum (p.name, p.name, p.name, enum_type_for (fi.type_name, &p), (gint) p.def, fl); break;
    }
    g_object_class_install_property (oc, PROP_FIRST + i, ps);
  }

  if (is_bayer_plugin_factory (fi.factory)) {
    GstCaps *raw = raw_caps_for (&fi);
    GstCaps *bayer = with_hbm_feature (gst_caps_from_string ("video/x-bayer,format=(string){bggr,grbg,gbrg,rggb},"
            "width=(int)[1,MAX],height=(int)[1,MAX],framerate=(fraction)[0/1,MAX]"));
    gboolean to_rgb = !strcmp (fi.factory, "bayer2rgb");
    gst_element_class_add_pad_template (ec, gst_pad_template_new ("src", GST_PAD_SRC, GST_PAD_ALWAYS, to_rgb ? raw : bayer));
    gst_element_class_add_pad_template (ec, gst_pad_template_new ("sink", GST_PAD_SINK, GST_PAD_ALWAYS, to_rgb ? bayer : raw));
    bc->transform_caps = GST_DEBUG_FUNCPTR (b200vf_bayer_transform_caps);
    bc->get_unit_size = GST_DEBUG_FUNCPTR (b200vf_bayer_get_unit_size);
    bc->set_caps = GST_DEBUG_FUNCPTR (b200vf_bayer_set_caps);
    bc->transform = GST_DEBUG_FUNCPTR (b200vf_bayer_transform);
    gst_caps_unref (raw);
    gst_caps_unref (bayer);
  } else {
    GstVideoFilterClass *vc = GST_VIDEO_FILTER_CLASS (g_class);
    GstCaps *raw = raw_caps_for (&fi);
    gst_element_class_add_pad_template (ec, gst_pad_template_new ("src", GST_PAD_SRC, GST_PAD_ALWAYS, raw));
    gst_element_class_add_pad_template (ec, gst_pad_template_new ("sink", GST_PAD_SINK, GST_PAD_ALWAYS, raw));
    gst_caps_unref (raw);
    vc->set_info = GST_DEBUG_FUNCPTR (b200vf_set_info);
    /* (GstVideoFilter's class_init installed its own transform / transform_ip, which map to system memory first) */
    if (fi.in_place) bc->transform_ip = GST_DEBUG_FUNCPTR (b200vf_vf_transform_ip);
    else bc->transform = GST_DEBUG_FUNCPTR (b200vf_vf_transform);
  }
  bc->propose_allocation = GST_DEBUG_FUNCPTR (b200vf_propose_allocation);
  bc->decide_allocation = GST_DEBUG_FUNCPTR (b200vf_decide_allocation);
  bc->start = GST_DEBUG_FUNCPTR (b200vf_start);
  bc->stop = GST_DEBUG_FUNCPTR (b200vf_stop);
  bc->before_transform = GST_DEBUG_FUNCPTR (b200vf_before_transform);
}

static void
b200vf_instance_init (GTypeInstance * instance, gpointer g_class)
{
  GstB200vf *self = B200VF (instance);
  self->device = 0;
  /* a device-less mirror holds the property values until start() binds the GPU */
  b200vf_element_factory_make (NULL, ((GstB200vfClass *) g_class)->factory, &self->el);
  if (is_bayer_plugin_factory (((GstB200vfClass *) g_class)->factory))
    gst_base_transform_set_in_place (GST_BASE_TRANSFORM (instance), TRUE);     /* gstbayer2rgb.c:209 */
}

/* abstract intermediate types keep the reference's GType hierarchy (gstgeometrictransform.c:406-432) */
static GType
abstract_type (const gchar * name, GType parent)
{
  GType t = g_type_from_name (name);
  if (!t) {
    GTypeInfo info = { sizeof (GstB200vfClass), NULL, NULL, NULL, NULL, NULL, sizeof (GstB200vf), 0, NULL, NULL };
    t = g_type_register_static (parent, name, &info, G_TYPE_FLAG_ABSTRACT);
  }
  return t;
}

static gboolean
plugin_init (GstPlugin * plugin)
{
  GST_DEBUG_CATEGORY_INIT (b200vf_debug, "b200vf", 0, "B200 video filters");
  gboolean ok = TRUE;
  for (gint i = 0; i < b200vf_factory_count (); i++) {
    b200vf_factory_info fi;
    if (b200vf_factory_get (i, &fi) != B200VF_OK || strcmp (fi.plugin, STR (B200VF_PLUGIN))) continue;
    GType parent = GST_TYPE_VIDEO_FILTER;
    if (!strcmp (fi.parent_type_name, "GstBaseTransform")) parent = GST_TYPE_BASE_TRANSFORM;
    else if (!strcmp (fi.parent_type_name, "GstGeometricTransform")) parent = abstract_type ("GstGeometricTransform", GST_TYPE_VIDEO_FILTER);
    else if (!strcmp (fi.parent_type_name, "GstCircleGeometricTransform"))
      parent = abstract_type ("GstCircleGeometricTransform", abstract_type ("GstGeometricTransform", GST_TYPE_VIDEO_FILTER));
    GTypeInfo info = { sizeof (GstB200vfClass), NULL, NULL, b200vf_class_init, NULL, fi.factory, sizeof (GstB200vf), 0,
      b200vf_instance_init, NULL };
    GType t = g_type_register_static (parent, fi.type_name, &info, 0);
    ok &= gst_element_register (plugin, fi.factory, GST_RANK_NONE, t);
  }
  return ok;
}

/* one level of indirection so that B200VF_PLUGIN is expanded before GST_PLUGIN_DEFINE pastes it */
#define B200VF_DEFINE_PLUGIN(name) \
  GST_PLUGIN_DEFINE (GST_VERSION_MAJOR, GST_VERSION_MINOR, name, "B200-native " STR (name) " (sm_100a kernels behind the stock element surface)", \
      plugin_init, VERSION, "LGPL", PACKAGE, "https://gstreamer.freedesktop.org")
B200VF_DEFINE_PLUGIN (B200VF_PLUGIN)
