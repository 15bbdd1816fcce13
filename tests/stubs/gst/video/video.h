/* stub: the part of <gst/video/video.h> the shells use */
#ifndef STUB_GST_VIDEO_H
#define STUB_GST_VIDEO_H
#include <gst/gst.h>
#define GST_VIDEO_MAX_PLANES 4
typedef enum { GST_VIDEO_FORMAT_UNKNOWN = 0 } GstVideoFormat;
typedef enum { GST_VIDEO_FRAME_FLAG_NONE = 0 } GstVideoFrameFlags;
typedef struct _GstVideoFormatInfo { GstVideoFormat format; const gchar *name; guint n_components; guint n_planes; } GstVideoFormatInfo;
typedef struct _GstVideoInfo {
  const GstVideoFormatInfo *finfo; gint interlace_mode; guint flags; gint width; gint height; gsize size; gint views;
  gsize offset[GST_VIDEO_MAX_PLANES]; gint stride[GST_VIDEO_MAX_PLANES];
} GstVideoInfo;
#define GST_VIDEO_INFO_FORMAT(i) ((i)->finfo->format)
#define GST_VIDEO_INFO_WIDTH(i) ((i)->width)
#define GST_VIDEO_INFO_HEIGHT(i) ((i)->height)
#define GST_VIDEO_INFO_SIZE(i) ((i)->size)
#define GST_VIDEO_INFO_N_PLANES(i) ((i)->finfo->n_planes)
typedef struct _GstVideoFrame { GstVideoInfo info; GstVideoFrameFlags flags; GstBuffer *buffer; gpointer meta; gint id; gpointer data[GST_VIDEO_MAX_PLANES]; GstMapInfo map[GST_VIDEO_MAX_PLANES]; } GstVideoFrame;
#define GST_VIDEO_FRAME_N_PLANES(f) (GST_VIDEO_INFO_N_PLANES (&(f)->info))
#define GST_VIDEO_FRAME_PLANE_DATA(f, p) ((f)->data[p])
#define GST_VIDEO_FRAME_PLANE_STRIDE(f, p) ((f)->info.stride[p])
typedef struct _GstVideoMeta {
  gpointer meta[2]; GstBuffer *buffer; GstVideoFrameFlags flags; GstVideoFormat format; gint id; guint width; guint height; guint n_planes;
  gsize offset[GST_VIDEO_MAX_PLANES]; gint stride[GST_VIDEO_MAX_PLANES];
} GstVideoMeta;
GType gst_video_meta_api_get_type (void);
#define GST_VIDEO_META_API_TYPE (gst_video_meta_api_get_type ())
#define GST_BUFFER_POOL_OPTION_VIDEO_META "GstBufferPoolOptionVideoMeta"
#define GST_VIDEO_CAPS_MAKE(format) "video/x-raw, format = (string) " format ", width = (int) [ 1, max ], height = (int) [ 1, max ], framerate = (fraction) [ 0, max ]"
gboolean gst_video_info_from_caps (GstVideoInfo * info, const GstCaps * caps);
const gchar *gst_video_format_to_string (GstVideoFormat format);
gboolean gst_video_frame_map (GstVideoFrame * frame, const GstVideoInfo * info, GstBuffer * buffer, GstMapFlags flags);
void gst_video_frame_unmap (GstVideoFrame * frame);
GstVideoMeta *gst_buffer_get_video_meta (GstBuffer * buffer);
GstVideoMeta *gst_buffer_add_video_meta_full (GstBuffer * buffer, GstVideoFrameFlags flags, GstVideoFormat format, guint width, guint height,
    guint n_planes, const gsize offset[GST_VIDEO_MAX_PLANES], const gint stride[GST_VIDEO_MAX_PLANES]);
GstEvent *gst_video_event_new_downstream_force_key_unit (GstClockTime timestamp, GstClockTime stream_time, GstClockTime running_time,
    gboolean all_headers, guint count);
#endif
