/* stub: GstVideoFilter (vfunc list as in tools/element-templates/videofilter of the reference) */
#ifndef STUB_GST_VIDEO_FILTER_H
#define STUB_GST_VIDEO_FILTER_H
#include <gst/base/gstbasetransform.h>
#include <gst/video/video.h>
typedef struct _GstVideoFilter { GstBaseTransform element; gboolean negotiated; GstVideoInfo in_info; GstVideoInfo out_info; gpointer _gst_reserved[4]; } GstVideoFilter;
typedef struct _GstVideoFilterClass {
  GstBaseTransformClass parent_class;
  gboolean (*set_info) (GstVideoFilter * filter, GstCaps * incaps, GstVideoInfo * in_info, GstCaps * outcaps, GstVideoInfo * out_info);
  GstFlowReturn (*transform_frame) (GstVideoFilter * filter, GstVideoFrame * inframe, GstVideoFrame * outframe);
  GstFlowReturn (*transform_frame_ip) (GstVideoFilter * trans, GstVideoFrame * frame);
  gpointer _gst_reserved[4];
} GstVideoFilterClass;
GType gst_video_filter_get_type (void);
#define GST_TYPE_VIDEO_FILTER (gst_video_filter_get_type ())
#define GST_VIDEO_FILTER_CAST(obj) ((GstVideoFilter *) (obj))
#define GST_VIDEO_FILTER_CLASS(klass) ((GstVideoFilterClass *) (klass))
#endif
