/* stub: GstBaseTransform as the shells use it (vfunc list as in tools/element-templates/basetransform of the reference) */
#ifndef STUB_GST_BASE_TRANSFORM_H
#define STUB_GST_BASE_TRANSFORM_H
#include <gst/gst.h>
typedef struct _GstBaseTransform { GstElement element; GstPad *sinkpad; GstPad *srcpad; gboolean have_segment; GstSegment segment; GstCaps *queued_buf; gpointer priv; } GstBaseTransform;
typedef struct _GstBaseTransformClass GstBaseTransformClass;
struct _GstBaseTransformClass {
  GstElementClass parent_class;
  gboolean passthrough_on_same_caps;
  gboolean transform_ip_on_passthrough;
  GstCaps *(*transform_caps) (GstBaseTransform * trans, GstPadDirection direction, GstCaps * caps, GstCaps * filter);
  GstCaps *(*fixate_caps) (GstBaseTransform * trans, GstPadDirection direction, GstCaps * caps, GstCaps * othercaps);
  gboolean (*accept_caps) (GstBaseTransform * trans, GstPadDirection direction, GstCaps * caps);
  gboolean (*set_caps) (GstBaseTransform * trans, GstCaps * incaps, GstCaps * outcaps);
  gboolean (*query) (GstBaseTransform * trans, GstPadDirection direction, GstQuery * query);
  gboolean (*decide_allocation) (GstBaseTransform * trans, GstQuery * query);
  gboolean (*filter_meta) (GstBaseTransform * trans, GstQuery * query, GType api, const GstStructure * params);
  gboolean (*propose_allocation) (GstBaseTransform * trans, GstQuery * decide_query, GstQuery * query);
  gboolean (*transform_size) (GstBaseTransform * trans, GstPadDirection direction, GstCaps * caps, gsize size, GstCaps * othercaps, gsize * othersize);
  gboolean (*get_unit_size) (GstBaseTransform * trans, GstCaps * caps, gsize * size);
  gboolean (*start) (GstBaseTransform * trans);
  gboolean (*stop) (GstBaseTransform * trans);
  gboolean (*sink_event) (GstBaseTransform * trans, GstEvent * event);
  gboolean (*src_event) (GstBaseTransform * trans, GstEvent * event);
  GstFlowReturn (*prepare_output_buffer) (GstBaseTransform * trans, GstBuffer * input, GstBuffer ** outbuf);
  gboolean (*copy_metadata) (GstBaseTransform * trans, GstBuffer * input, GstBuffer * outbuf);
  gboolean (*transform_meta) (GstBaseTransform * trans, GstBuffer * outbuf, gpointer meta, GstBuffer * inbuf);
  void (*before_transform) (GstBaseTransform * trans, GstBuffer * buffer);
  GstFlowReturn (*transform) (GstBaseTransform * trans, GstBuffer * inbuf, GstBuffer * outbuf);
  GstFlowReturn (*transform_ip) (GstBaseTransform * trans, GstBuffer * buf);
};
GType gst_base_transform_get_type (void);
#define GST_TYPE_BASE_TRANSFORM (gst_base_transform_get_type ())
#define GST_BASE_TRANSFORM(obj) ((GstBaseTransform *) (obj))
#define GST_BASE_TRANSFORM_CLASS(klass) ((GstBaseTransformClass *) (klass))
#define GST_BASE_TRANSFORM_SRC_PAD(obj) (GST_BASE_TRANSFORM (obj)->srcpad)
void gst_base_transform_set_in_place (GstBaseTransform * trans, gboolean in_place);
#endif
