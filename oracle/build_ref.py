#!/usr/bin/env python3
"""Build oracle/_ref/libgstbad_ref.so from the reference's OWN sources.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or
executed by the product path (gst-plugins-bad_b200/); only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
use it, and only as the checker / the CPU baseline.

What it does: the reference's element shells need
GLib/GObject/GStreamer/ORC, none of which exist in this image, but the
per-pixel inner loops are plain C.  This script
  * compiles the ORC C backups `gst/bayer/gstbayerorc-dist.c` and
    `gst/gaudieffects/gstgaudieffectsorc-dist.c` unmodified with -DDISABLE_ORC
    (the path the reference's own meson build takes without ORC,
    /root/reference/meson.build:409-416), and
  * extracts the `static` scalar loops BY FUNCTION NAME at build time from the
    files where they lie under /root/reference, wraps each in a translation
    unit made of a typedef shim (oracle/shim/glib.h) + a plain-C `ref_*` entry
    point, and compiles that.
Generated translation units live in a temporary directory and are deleted; only
the shared object is written, to oracle/_ref/ (git-ignored; it travels to the
GPU box with the snapshot).  No reference source is copied into the repository.

Flags: -O2 -ffp-contract=off, generic x86-64 (no -march=native) so that fp32
results of gaussianblur do not depend on FMA contraction.
"""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("B200VF_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libgstbad_ref.so")
CFLAGS = ["-std=gnu99", "-O2", "-fPIC", "-ffp-contract=off", "-DDISABLE_ORC",
          "-DGST_DISABLE_GST_DEBUG", "-w", "-I", os.path.join(HERE, "shim")]


def read(rel):
    with open(os.path.join(REF, rel), "r", encoding="utf-8", errors="replace") as f:
        return f.read().split("\n")


def func(rel, name):
    """Return the text of the C function `name` defined in reference file `rel`
    (GStreamer style: return type on its own line, name at column 0, closing
    brace at column 0)."""
    ls = read(rel)
    pat = re.compile(r"^%s \(" % re.escape(name))
    for i, l in enumerate(ls):
        if pat.match(l) and not ls[i].rstrip().endswith(";"):
            # declaration lines end with ';' within a few lines; definitions reach '{'
            j = i
            while j < len(ls) and "{" not in ls[j] and ";" not in ls[j]:
                j += 1
            if j < len(ls) and ls[j].strip() == "{" or (j < len(ls) and ls[j].rstrip().endswith("{")):
                start = i - 1
                while start > 0 and ls[start - 1].strip() and not ls[start - 1].startswith(("}", "#", "/*", " *")):
                    start -= 1
                end = j
                while ls[end] != "}":
                    end += 1
                return "\n".join(ls[start:end + 1]) + "\n"
    raise RuntimeError("function %s not found in %s" % (name, rel))


def rng(rel, a, b):
    """Lines a..b (1-based, inclusive) of a reference file."""
    return "\n".join(read(rel)[a - 1:b]) + "\n"


def between(rel, start_pat, end_pat, include_end=True):
    ls = read(rel)
    s = next(i for i, l in enumerate(ls) if re.search(start_pat, l))
    e = next(i for i in range(s, len(ls)) if re.search(end_pat, ls[i]))
    return "\n".join(ls[s:e + (1 if include_end else 0)]) + "\n"


def unstatic(text, name):
    return text


TUS = {}

# ---------------------------------------------------------------- bayer2rgb
GB = "gst/bayer/"
TUS["ref_bayer2rgb.c"] = lambda: (
    '#include <glib.h>\n#include "gstbayerorc-dist.h"\n'
    "typedef struct { int width, height, r_off, g_off, b_off, format; } GstBayer2RGB;\n"
    + between(GB + "gstbayer2rgb.c", r"^enum$", r"^};")           # format enum :95-101
    + func(GB + "gstbayer2rgb.c", "gst_bayer2rgb_split_and_upsample_horiz")
    + between(GB + "gstbayer2rgb.c", r"^typedef void \(\*process_func\)", r"int n\);")
    + func(GB + "gstbayer2rgb.c", "gst_bayer2rgb_process")
    + """
void ref_bayer2rgb (uint8_t *dest, int dest_stride, uint8_t *src, int src_stride,
    int width, int height, int format, int r_off, int g_off, int b_off)
{
  GstBayer2RGB b = { width, height, r_off, g_off, b_off, format };
  gst_bayer2rgb_process (&b, dest, dest_stride, src, src_stride);
}
""")


TUS["ref_rgb2bayer.c"] = lambda: (
    "#include <glib.h>\n"
    "void ref_rgb2bayer (guint8 *dest, guint8 *src, int src_stride, int width, int height, int format)\n"
    "{\n  int i, j;\n  struct { int format; } rb = { format }, *rgb2bayer = &rb;\n"
    "  struct { struct { int stride[1]; } info; } frame; frame.info.stride[0] = src_stride;\n"
    # the per-pixel select loop, gstrgb2bayer.c:254-267
    + between(GB + "gstrgb2bayer.c", r"^  for \(j = 0; j < height; j\+\+\) \{", r"^  }$")
    + "}\n")

# ------------------------------------------------------------- gaudieffects
GG = "gst/gaudieffects/"
TUS["ref_gaussblur.c"] = lambda: (
    "#include <glib.h>\n"
    "typedef struct { gint width, height, stride; float cur_sigma, sigma; int windowsize;\n"
    "  float *kernel; float *kernel_sum; float *tempim; } GstGaussianBlur;\n"
    "static gboolean make_gaussian_kernel (GstGaussianBlur * gb, float sigma);\n"
    + func(GG + "gstgaussblur.c", "blur_row_x")
    + func(GG + "gstgaussblur.c", "gaussian_smooth")
    + func(GG + "gstgaussblur.c", "make_gaussian_kernel")
    + """
/* kernel/kernel_sum out arrays must hold >= 2*ceil(2.5*|sigma|)+1 floats */
int ref_gauss_kernel (float sigma, float *kernel, float *kernel_sum)
{
  GstGaussianBlur gb; memset (&gb, 0, sizeof gb);
  if (!make_gaussian_kernel (&gb, sigma)) return -1;
  memcpy (kernel, gb.kernel, sizeof (float) * gb.windowsize);
  memcpy (kernel_sum, gb.kernel_sum, sizeof (float) * gb.windowsize);
  free (gb.kernel); free (gb.kernel_sum);
  return gb.windowsize;
}
/* image/out_image are the COMP_DATA(frame,0) pointers (plane + p0); the caller
 * has already done the gst_video_frame_copy (gstgaussblur.c:252). sigma is the
 * float the element snapshots from its double property (:229). */
int ref_gaussblur (guint8 *image, guint8 *out_image, int width, int height, int stride, float sigma)
{
  GstGaussianBlur gb; memset (&gb, 0, sizeof gb);
  gb.width = width; gb.height = height; gb.stride = stride;
  gb.sigma = gb.cur_sigma = sigma;
  if (!make_gaussian_kernel (&gb, sigma)) return -1;
  gb.tempim = malloc (sizeof (float) * (size_t) stride * height + 64);
  if (gb.sigma != 0.0) gaussian_smooth (&gb, image, out_image);
  free (gb.tempim); free (gb.kernel); free (gb.kernel_sum);
  return 0;
}
""")


def point_tu(cfile, extra_pre, call_sig, call_body, extra_funcs=()):
    def gen():
        s = "#include <glib.h>\n" + extra_pre
        for fn in extra_funcs:
            s += func(GG + cfile, fn)
        s += func(GG + cfile, "transform")
        s += call_sig + "\n{\n" + call_body + "\n}\n"
        return s
    return gen


TUS["ref_burn.c"] = lambda: (
    '#include <glib.h>\n#include "gstgaudieffectsorc-dist.h"\n'
    "void ref_burn (guint32 *dest, const guint32 *src, int adjustment, int n)\n"
    "{ gaudi_orc_burn (dest, src, adjustment, n); }\n")
TUS["ref_dodge.c"] = point_tu("gstdodge.c", "",
    "void ref_dodge (guint32 *src, guint32 *dest, int n)", "  transform (src, dest, n);")
TUS["ref_chromium.c"] = point_tu("gstchromium.c",
    between("gst/gaudieffects/gstchromium.c", r"^const float pi", r"^gint cosTable\[")
    + "static gint cos_from_table (int angle);\n",
    "void ref_chromium (guint32 *src, guint32 *dest, int n, int edge_a, int edge_b)",
    "  setup_cos_table ();\n  transform (src, dest, n, edge_a, edge_b);",
    extra_funcs=("setup_cos_table", "abs_int", "cos_from_table"))
TUS["ref_dilate.c"] = point_tu("gstdilate.c", "",
    "void ref_dilate (guint32 *src, guint32 *dest, int n, int width, int height, int erode)",
    "  transform (src, dest, n, width, height, erode);", extra_funcs=("get_luminance",))
TUS["ref_exclusion.c"] = point_tu("gstexclusion.c", "",
    "void ref_exclusion (guint32 *src, guint32 *dest, int n, int factor)",
    "  transform (src, dest, n, factor);")
TUS["ref_solarize.c"] = point_tu("gstsolarize.c", "",
    "void ref_solarize (guint32 *src, guint32 *dest, int n, int threshold, int start, int end)",
    "  transform (src, dest, n, threshold, start, end);")

# ------------------------------------------------------------- coloreffects
GC = "gst/coloreffects/"
FRAME_SHIM = """
typedef struct { guint8 *data; int width, height, stride, pstride; int poffset[4]; } GstVideoFrame;
#define GST_VIDEO_FRAME_PLANE_DATA(f,p) ((f)->data)
#define GST_VIDEO_FRAME_COMP_POFFSET(f,c) ((f)->poffset[c])
#define GST_VIDEO_FRAME_WIDTH(f) ((f)->width)
#define GST_VIDEO_FRAME_HEIGHT(f) ((f)->height)
#define GST_VIDEO_FRAME_PLANE_STRIDE(f,p) ((f)->stride)
#define GST_VIDEO_FRAME_COMP_PSTRIDE(f,c) ((f)->pstride)
"""
TUS["ref_coloreffects.c"] = lambda: (
    "#include <glib.h>\n" + FRAME_SHIM
    + "typedef struct { const guint8 *table; gboolean map_luma; } GstColorEffects;\n"
    # the five 256x3 tables + the two cog matrices + APPLY_MATRIX (:116-301)
    + between(GC + "gstcoloreffects.c", r"^/\* 256 \* 3 RGB data \*/", r"^#define APPLY_MATRIX", False)
    + between(GC + "gstcoloreffects.c", r"^#define APPLY_MATRIX", r">> 8\)")
    + func(GC + "gstcoloreffects.c", "gst_color_effects_transform_rgb")
    + func(GC + "gstcoloreffects.c", "gst_color_effects_transform_ayuv")
    + """
/* preset numbering = GstColorEffectsPreset (gstcoloreffects.c:74-96):
 * 0 none, 1 heat, 2 sepia, 3 xray, 4 xpro, 5 yellowblue; mapping to
 * (table, map_luma) as in set_property (:503-548). */
static int preset_lookup (int preset, const guint8 **t, int *map_luma)
{
  switch (preset) {
    case 1: *t = heat_table; *map_luma = 1; return 1;
    case 2: *t = sepia_table; *map_luma = 1; return 1;
    case 3: *t = xray_table; *map_luma = 1; return 1;
    case 4: *t = xpro_table; *map_luma = 0; return 1;
    case 5: *t = yellowblue_table; *map_luma = 0; return 1;
    default: *t = NULL; *map_luma = 0; return 0;
  }
}
int ref_coloreffects_table (int preset, guint8 *out768, int *map_luma)
{
  const guint8 *t;
  if (!preset_lookup (preset, &t, map_luma)) return 0;
  memcpy (out768, t, 768);
  return 1;
}
void ref_coloreffects (guint8 *data, int width, int height, int stride, int pstride,
    int o0, int o1, int o2, int preset, int is_ayuv)
{
  GstVideoFrame f = { data, width, height, stride, pstride, { o0, o1, o2, 0 } };
  GstColorEffects ce; int ml;
  if (!preset_lookup (preset, &ce.table, &ml)) return;   /* none => no-op (:488-490) */
  ce.map_luma = ml;
  if (is_ayuv) gst_color_effects_transform_ayuv (&ce, &f);
  else gst_color_effects_transform_rgb (&ce, &f);
}
""")
TUS["ref_chromahold.c"] = lambda: (
    "#include <glib.h>\n" + FRAME_SHIM
    + "typedef struct { gint tolerance; gint hue; } GstChromaHold;\n"
    + func(GC + "gstchromahold.c", "rgb_to_hue")
    + func(GC + "gstchromahold.c", "hue_dist")
    + func(GC + "gstchromahold.c", "gst_chroma_hold_process_xrgb")
    + """
/* p3 = COMP_POFFSET(frame,3) (alpha/x), p0..p2 = R,G,B byte offsets */
void ref_chromahold (guint8 *data, int width, int height, int stride,
    int p0, int p1, int p2, int p3, int target_r, int target_g, int target_b, int tolerance)
{
  GstVideoFrame f = { data, width, height, stride, 4, { p0, p1, p2, p3 } };
  GstChromaHold ch;
  ch.tolerance = tolerance;
  ch.hue = rgb_to_hue (target_r, target_g, target_b);   /* init_params :362-366 */
  gst_chroma_hold_process_xrgb (&f, width, height, &ch);
}
int ref_rgb_to_hue (int r, int g, int b) { return rgb_to_hue (r, g, b); }
""")

# ------------------------------------------------------- geometrictransform
GT = "gst/geometrictransform/"
GT_SHIM = """
#include <glib.h>
typedef struct _GstGeometricTransform GstGeometricTransform;
typedef struct _GstGMNoise GstGMNoise;
struct _GstGeometricTransform { gint width, height; gint pixel_stride; gint row_stride;
  gboolean precalc_map; gboolean needs_remap; gint off_edge_pixels; gdouble *map; };
typedef struct { GstGeometricTransform element; gdouble x_center, y_center, radius;
  gdouble precalc_x_center, precalc_y_center, precalc_radius, precalc_radius2; } GstCircleGeometricTransform;
#define GST_GEOMETRIC_TRANSFORM_CAST(o) ((GstGeometricTransform *)(o))
#define GST_CIRCLE_GEOMETRIC_TRANSFORM_CAST(o) ((GstCircleGeometricTransform *)(o))
enum { GST_GT_OFF_EDGES_PIXELS_IGNORE = 0, GST_GT_OFF_EDGES_PIXELS_CLAMP, GST_GT_OFF_EDGES_PIXELS_WRAP };
gdouble gst_gm_mod_float (gdouble a, gdouble b);
gdouble gst_gm_triangle (gdouble x);
gdouble gst_gm_smoothstep (gdouble edge0, gdouble edge1, gdouble x);
gdouble gst_gm_noise_2 (GstGMNoise * noise, gdouble x, gdouble y);
typedef gboolean (*ref_map_func) (GstGeometricTransform * gt, gint x, gint y, gdouble * in_x, gdouble * in_y);
"""
TUS["ref_gt_math.c"] = lambda: (
    GT_SHIM + "#define g_random_int() ((guint32) random ())\n"
    + between(GT + "geometricmath.c", r"^#define N +0x1000", r"^#define BM")
    + between(GT + "geometricmath.c", r"^struct _GstGMNoise", r"^};")
    + func(GT + "geometricmath.c", "normalize_2")
    + func(GT + "geometricmath.c", "gst_gm_noise_new")
    + func(GT + "geometricmath.c", "s_curve") + func(GT + "geometricmath.c", "lerp")
    + func(GT + "geometricmath.c", "gst_gm_noise_2")
    + func(GT + "geometricmath.c", "gst_gm_mod_float")
    + func(GT + "geometricmath.c", "gst_gm_triangle")
    + func(GT + "geometricmath.c", "gst_gm_smoothstep"))

TUS["ref_gt_base.c"] = lambda: (
    GT_SHIM
    + func(GT + "gstgeometrictransform.c", "gst_geometric_transform_do_map")
    + """
/* gst_geometric_transform_generate_map (:80-128) + the precalc branch of
 * transform_frame (:244-273), driven through the extracted do_map. */
int ref_gt_generate_map (ref_map_func map_func, GstGeometricTransform *gt, double *map)
{
  gint x, y; gdouble in_x, in_y; gdouble *ptr = map;
  for (y = 0; y < gt->height; y++)
    for (x = 0; x < gt->width; x++) {
      if (!map_func (gt, x, y, &in_x, &in_y)) return 0;
      ptr[0] = in_x; ptr[1] = in_y; ptr += 2;
    }
  return 1;
}
void ref_gt_apply_map (GstGeometricTransform *gt, const double *map, guint8 *in_data,
    guint8 *out_data, size_t out_size, int is_ayuv)
{
  gint x, y; size_t i; const gdouble *ptr = map;
  if (is_ayuv) {
    for (i = 0; i + 4 <= out_size; i += 4) {   /* GST_WRITE_UINT32_BE (.., 0xff108080) :244-250 */
      out_data[i] = 0xff; out_data[i + 1] = 0x10; out_data[i + 2] = 0x80; out_data[i + 3] = 0x80;
    }
  } else memset (out_data, 0, out_size);
  for (y = 0; y < gt->height; y++)
    for (x = 0; x < gt->width; x++) {
      gst_geometric_transform_do_map (gt, in_data, out_data, x, y, ptr[0], ptr[1]);
      ptr += 2;
    }
}
""")

# element table: name -> (file, struct text, cast macro, map func, prepare func or None,
#                         init code for defaults, property setter code)
GT_ELEMENTS = {}


def gt_element(name, cfile, struct_body, cast, mapf, prepare=None, is_circle=False,
               extra_funcs=(), extra_pre="", props=()):
    GT_ELEMENTS[name] = dict(cfile=cfile, struct_body=struct_body, cast=cast, mapf=mapf,
                             prepare=prepare, is_circle=is_circle, extra_funcs=extra_funcs,
                             extra_pre=extra_pre, props=props)


gt_element("fisheye", "gstfisheye.c", "GstGeometricTransform element;", "GST_FISHEYE_CAST", "fisheye_map")
C = "GstCircleGeometricTransform element;"
Gt = "GstGeometricTransform element;"
gt_element("bulge", "gstbulge.c", C + " gdouble zoom;", "GST_BULGE_CAST", "bulge_map", is_circle=True,
           props=[("zoom", "zoom", "gdouble")])
gt_element("circle", "gstcircle.c", C + " gdouble angle; gdouble spread_angle; gint height;", "GST_CIRCLE_CAST", "circle_map",
           is_circle=True, props=[("angle", "angle", "gdouble"), ("spread_angle", "spread_angle", "gdouble"), ("height", "height", "gint")])
gt_element("kaleidoscope", "gstkaleidoscope.c", C + " gdouble angle; gdouble angle2; gint sides;", "GST_KALEIDOSCOPE_CAST",
           "kaleidoscope_map", is_circle=True,
           props=[("angle", "angle", "gdouble"), ("angle2", "angle2", "gdouble"), ("sides", "sides", "gint")])
gt_element("pinch", "gstpinch.c", C + " gdouble intensity;", "GST_PINCH_CAST", "pinch_map", is_circle=True,
           props=[("intensity", "intensity", "gdouble")])
gt_element("rotate", "gstrotate.c", Gt + " gdouble angle;", "GST_ROTATE_CAST", "rotate_map", props=[("angle", "angle", "gdouble")])
gt_element("sphere", "gstsphere.c", C + " gdouble refraction;", "GST_SPHERE_CAST", "sphere_map", is_circle=True,
           props=[("refraction", "refraction", "gdouble")])
gt_element("twirl", "gsttwirl.c", C + " gdouble angle;", "GST_TWIRL_CAST", "twirl_map", is_circle=True,
           props=[("angle", "angle", "gdouble")])
gt_element("waterripple", "gstwaterripple.c", C + " gdouble phase; gdouble amplitude; gdouble wavelength;", "GST_WATER_RIPPLE_CAST",
           "water_ripple_map", is_circle=True,
           props=[("phase", "phase", "gdouble"), ("amplitude", "amplitude", "gdouble"), ("wavelength", "wavelength", "gdouble")])
gt_element("stretch", "gststretch.c", C + " gdouble intensity;", "GST_STRETCH_CAST", "stretch_map", is_circle=True,
           extra_pre="#define MAX_SHRINK_AMOUNT 3.0\n", props=[("intensity", "intensity", "gdouble")])
gt_element("tunnel", "gsttunnel.c", C, "GST_TUNNEL_CAST", "tunnel_map", is_circle=True)
gt_element("square", "gstsquare.c", Gt + " gdouble width, height; gdouble zoom;", "GST_SQUARE_CAST", "square_map",
           props=[("width", "width", "gdouble"), ("height", "height", "gdouble"), ("zoom", "zoom", "gdouble")])
gt_element("mirror", "gstmirror.c", Gt + " gint mode;", "GST_MIRROR_CAST", "mirror_map",
           extra_pre="enum { GST_MIRROR_MODE_LEFT, GST_MIRROR_MODE_RIGHT, GST_MIRROR_MODE_TOP, GST_MIRROR_MODE_BOTTOM };\n"
                     "#define g_assert_not_reached() do { } while (0)\n",
           props=[("mode", "mode", "gint")])
gt_element("perspective", "gstperspective.c", Gt + " gdouble matrix[9];", "GST_PERSPECTIVE_CAST", "perspective_map",
           props=[("matrix_%d" % i, "matrix[%d]" % i, "gdouble") for i in range(9)])

CIRCLE_PRECALC = None


def gt_tu(name):
    e = GT_ELEMENTS[name]

    def gen():
        s = GT_SHIM
        s += "typedef struct { %s } RefElem_%s;\n" % (e["struct_body"], name)
        tname = {"waterripple": "GstWaterRipple"}.get(name, "Gst" + name.capitalize())
        s += "typedef RefElem_%s %s;\n" % (name, tname)
        s += "#define %s(o) ((RefElem_%s *)(o))\n" % (e["cast"], name)
        s += e["extra_pre"]
        for fn in e["extra_funcs"]:
            s += func(GT + e["cfile"], fn)
        if e["prepare"]:
            s += func(GT + e["cfile"], e["prepare"])
        s += func(GT + e["cfile"], e["mapf"])
        if e["is_circle"]:
            s += func(GT + "gstcirclegeometrictransform.c", "circle_geometric_transform_precalc")
        s += "size_t ref_gt_%s_size (void) { return sizeof (RefElem_%s); }\n" % (name, name)
        s += "ref_map_func ref_gt_%s_map (void) { return %s; }\n" % (name, e["mapf"])
        prep = []
        if e["is_circle"]:
            prep.append("circle_geometric_transform_precalc ((GstGeometricTransform *) p);")
        if e["prepare"]:
            prep.append("%s ((GstGeometricTransform *) p);" % e["prepare"])
        s += "void ref_gt_%s_prepare (void *p) { %s }\n" % (name, " ".join(prep))
        # property setters by name: offsets exported so the python side can poke doubles/ints
        for pname, field, ctype in e["props"]:
            s += ("void ref_gt_%s_set_%s (void *p, double v) { ((RefElem_%s *) p)->%s = (%s) v; }\n"
                  % (name, pname, name, field, ctype))
        if e["is_circle"]:
            for pname in ("x_center", "y_center", "radius"):
                s += ("void ref_gt_%s_set_%s (void *p, double v) { ((GstCircleGeometricTransform *) p)->%s = v; }\n"
                      % (name, pname, pname))
        return s
    return gen


# ------------------------------------------------------------- videofilters
GV = "gst/videofilters/"
VF_FRAME_SHIM = """
typedef int GstFlowReturn;
#define GST_FLOW_OK 0
typedef struct { struct { int width, height; int stride[4]; int comp_w[4], comp_h[4]; } info; void *data[4]; } GstVideoFrame;
#define GST_VIDEO_FRAME_COMP_HEIGHT(f,c) ((f)->info.comp_h[c])
#define GST_VIDEO_FRAME_COMP_WIDTH(f,c) ((f)->info.comp_w[c])
"""
TUS["ref_zebrastripe.c"] = lambda: (
    "#include <glib.h>\n"
    "/* data0 = frame->data[0]; offset / y_position / pixel_stride as the format switch sets them (:219-240) */\n"
    "void ref_zebrastripe (guint8 *data0, int stride0, int width, int height, int threshold, int t,\n"
    "    int offset, int pixel_stride, int y_position)\n{\n  int i, j;\n"
    "  struct { void *data[1]; struct { int stride[1]; } info; } fr, *frame = &fr;\n"
    "  fr.data[0] = data0; fr.info.stride[0] = stride0;\n"
    # the per-pixel loop, gstzebrastripe.c:242-250
    + between(GV + "gstzebrastripe.c", r"^  for \(j = 0; j < height; j\+\+\) \{", r"^  }$")
    + "}\n"
    # y_threshold from the property (:150-151)
    + "int ref_zebrastripe_y_threshold (int threshold)\n{\n  struct { int threshold, y_threshold; } z, *zebrastripe = &z;\n"
    "  z.threshold = threshold;\n"
    + between(GV + "gstzebrastripe.c", r"^      zebrastripe->y_threshold =$", r"2\.19 \* zebrastripe->threshold\);")
    + "  return z.y_threshold;\n}\n")

TUS["ref_videodiff.c"] = lambda: (
    "#include <glib.h>\n" + VF_FRAME_SHIM
    + "typedef struct { int threshold; int t; } GstVideoDiff;\n"
    + func(GV + "gstvideodiff.c", "gst_video_diff_transform_frame_ip_planarY")
    + """
/* three planes each; comp_w/comp_h = GST_VIDEO_FRAME_COMP_WIDTH/HEIGHT of planes 1 and 2 */
void ref_videodiff (guint8 *out[3], guint8 *in[3], guint8 *old[3], const int stride[3], int width, int height,
    int cw, int ch, int threshold, int t)
{
  GstVideoDiff vd = { threshold, t };
  GstVideoFrame o, n, p;
  GstVideoFrame *fr[3] = { &o, &n, &p };
  guint8 **pl[3] = { out, in, old };
  for (int f = 0; f < 3; f++) {
    fr[f]->info.width = width; fr[f]->info.height = height;
    for (int k = 0; k < 3; k++) {
      fr[f]->data[k] = pl[f][k]; fr[f]->info.stride[k] = stride[k];
      fr[f]->info.comp_w[k] = k ? cw : width; fr[f]->info.comp_h[k] = k ? ch : height;
    }
  }
  gst_video_diff_transform_frame_ip_planarY (&vd, &o, &n, &p);
}
""")

TUS["ref_scenechange.c"] = lambda: (
    "#include <glib.h>\n"
    "void orc_sad_nxm_u8 (guint32 * a1, const guint8 * s1, int s1_stride, const guint8 * s2, int s2_stride, int n, int m);\n"
    "#define SC_N_DIFFS 5\n"
    "typedef struct { int n_diffs; double diffs[SC_N_DIFFS]; } GstSceneChange;\n"
    "/* get_frame_score, gstscenechange.c:141-155 */\n"
    "double ref_scenechange_score (const guint8 *a, int a_stride, const guint8 *b, int b_stride, int width, int height, guint32 *sad)\n"
    "{\n  guint32 score = 0;\n  orc_sad_nxm_u8 (&score, a, a_stride, b, b_stride, width, height);\n  if (sad) *sad = score;\n"
    "  return ((double) score) / (width * height);\n}\n"
    "/* the decision, :196-236; state = { n_diffs, diffs[5] } */\n"
    "int ref_scenechange_update (GstSceneChange *scenechange, double score)\n{\n"
    "  double score_min, score_max, threshold;\n  gboolean change;\n  int i;\n"
    + between(GV + "gstscenechange.c", r"^  memmove \(scenechange->diffs, scenechange->diffs \+ 1,$", r"^    scenechange->n_diffs = 0;$")
    + "  }\n  return change;\n}\n")


TUS["ref_videoanalyse.c"] = lambda: (
    "#include <glib.h>\n" + VF_FRAME_SHIM
    + "typedef struct { double luma_average, luma_variance; } GstVideoAnalyse;\n"
    + func("gst/videosignal/gstvideoanalyse.c", "gst_video_analyse_planar")
    + """
void ref_videoanalyse (guint8 *luma, int stride, int width, int height, double *average, double *variance)
{
  GstVideoAnalyse va; GstVideoFrame f;
  f.info.width = width; f.info.height = height; f.info.stride[0] = stride; f.data[0] = luma;
  gst_video_analyse_planar (&va, &f);
  *average = va.luma_average; *variance = va.luma_variance;
}
""")

# ------------------------------------------------------------- videosignal: simplevideomark / simplevideomarkdetect
VM_FRAME_SHIM = """
typedef int GstFlowReturn;
#define GST_FLOW_OK 0
typedef struct { struct { int width, height; } info; guint8 *data0; int stride0, pstride0; void *buffer; } GstVideoFrame;
#define GST_VIDEO_FRAME_COMP_STRIDE(f,c) ((f)->stride0)
#define GST_VIDEO_FRAME_COMP_PSTRIDE(f,c) ((f)->pstride0)
#define GST_VIDEO_FRAME_COMP_DATA(f,c) ((f)->data0)
#define GST_ERROR_OBJECT(...) do { } while (0)
#define G_GUINT64_CONSTANT(v) (v##ULL)
#define G_GUINT64_FORMAT "lu"
typedef void GstBuffer;
"""
TUS["ref_videomark.c"] = lambda: (
    "#include <glib.h>\n" + VM_FRAME_SHIM
    + "typedef struct { gint pattern_width, pattern_height, pattern_count, pattern_data_count; guint64 pattern_data;\n"
      "  gboolean enabled; gint left_offset, bottom_offset; } GstSimpleVideoMark;\n"
    + func("gst/videosignal/gstsimplevideomark.c", "gst_video_mark_draw_box")
    + func("gst/videosignal/gstsimplevideomark.c", "calculate_pw")
    + func("gst/videosignal/gstsimplevideomark.c", "gst_video_mark_yuv")
    + """
void ref_videomark (guint8 *data, int stride, int pstride, int width, int height, int pw, int ph, int pc, int pdc,
    guint64 pdata, int left, int bottom)
{
  GstSimpleVideoMark m = { pw, ph, pc, pdc, pdata, TRUE, left, bottom };
  GstVideoFrame f;
  f.info.width = width; f.info.height = height; f.data0 = data; f.stride0 = stride; f.pstride0 = pstride; f.buffer = 0;
  gst_video_mark_yuv (&m, &f);
}
""")

TUS["ref_videomarkdetect.c"] = lambda: (
    "#include <glib.h>\n" + VM_FRAME_SHIM
    + "typedef struct { gboolean message; gint pattern_width, pattern_height, pattern_count, pattern_data_count;\n"
      "  gdouble pattern_center, pattern_sensitivity; gint left_offset, bottom_offset; gboolean in_pattern; } GstSimpleVideoMarkDetect;\n"
      "static int ref_msgs; static guint64 ref_last_data;\n"
      "static void gst_video_detect_post_message (GstSimpleVideoMarkDetect *d, GstBuffer *b, guint64 data) { ref_msgs++; ref_last_data = data; }\n"
    + func("gst/videosignal/gstsimplevideomarkdetect.c", "gst_video_detect_calc_brightness")
    + func("gst/videosignal/gstsimplevideomarkdetect.c", "calculate_pw")
    + func("gst/videosignal/gstsimplevideomarkdetect.c", "gst_video_detect_yuv")
    + """
/* one frame; *in_pattern is the element's state across frames; returns the number of messages posted (0 or 1) and,
 * when one was, its "data" field */
int ref_videomarkdetect (guint8 *data, int stride, int pstride, int width, int height, int pw, int ph, int pc, int pdc,
    double center, double sensitivity, int left, int bottom, int *in_pattern, guint64 *msg_data)
{
  GstSimpleVideoMarkDetect d = { TRUE, pw, ph, pc, pdc, center, sensitivity, left, bottom, *in_pattern };
  GstVideoFrame f;
  f.info.width = width; f.info.height = height; f.data0 = data; f.stride0 = stride; f.pstride0 = pstride; f.buffer = 0;
  ref_msgs = 0; ref_last_data = 0;
  gst_video_detect_yuv (&d, &f);
  *in_pattern = d.in_pattern;
  *msg_data = ref_last_data;
  return ref_msgs;
}
""")

TUS["ref_smooth.c"] = lambda: (
    "#include <glib.h>\n"
    + func("gst/smooth/gstsmooth.c", "smooth_filter")
    + """
void ref_smooth_plane (guchar *dest, guchar *src, int width, int height, int stride, int dstride, int tolerance, int filtersize)
{
  smooth_filter (dest, src, width, height, stride, dstride, tolerance, filtersize);
}
""")


def register_gt_tus():
    for name in GT_ELEMENTS:
        TUS["ref_gt_%s.c" % name] = gt_tu(name)


def build(verbose=False):
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree %s is not present; oracle/_ref can only be (re)built "
                           "where /root/reference is mounted" % REF)
    register_gt_tus()
    os.makedirs(OUT_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="b200vf_ref_") as tmp:
        objs = []
        # ORC C backups, compiled where they lie
        for rel in ("gst/bayer/gstbayerorc-dist.c", "gst/gaudieffects/gstgaudieffectsorc-dist.c",
                    "gst/videofilters/gstscenechangeorc-dist.c"):
            o = os.path.join(tmp, os.path.basename(rel) + ".o")
            subprocess.check_call(["gcc"] + CFLAGS + ["-c", os.path.join(REF, rel), "-o", o])
            objs.append(o)
        inc = ["-I", os.path.join(REF, "gst/bayer"), "-I", os.path.join(REF, "gst/gaudieffects")]
        for fname, gen in TUS.items():
            src = os.path.join(tmp, fname)
            with open(src, "w") as f:
                f.write("/* GENERATED at build time from %s - not part of the repository */\n" % REF)
                f.write(gen())
            o = src + ".o"
            try:
                subprocess.check_call(["gcc"] + CFLAGS + inc + ["-c", src, "-o", o])
            except subprocess.CalledProcessError:
                if verbose:
                    sys.stderr.write(open(src).read())
                raise
            objs.append(o)
        subprocess.check_call(["gcc", "-shared", "-o", OUT] + objs + ["-lm"])
    return OUT


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
