"""numpy/ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Two back-ends with one call surface:
  * Port  - oracle/lib/liboracle_port.so, our C restatement (oracle/port/*.c);
  * Ref   - oracle/_ref/libgstbad_ref.so, the reference's own C compiled from
            /root/reference by oracle/build_ref.py (present where it was built;
            it travels to the GPU box with the snapshot).
`get(kind)` returns one of them; tests check Port == Ref == golden fixtures and
then use either as the checker for the CUDA path.  The product never imports
this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "lib", "liboracle_port.so")
REF_SO = os.path.join(HERE, "_ref", "libgstbad_ref.so")

BAYER_FORMATS = {"bggr": 0, "gbrg": 1, "grbg": 2, "rggb": 3}
# (r,g,b) byte offsets of the 8 src-caps formats (GST_VIDEO_INFO_COMP_OFFSET of
# gst-plugins-base's format table; consumed at gstbayer2rgb.c:269-271)
RGB_OFFSETS = {
    "RGBx": (0, 1, 2), "RGBA": (0, 1, 2), "BGRx": (2, 1, 0), "BGRA": (2, 1, 0),
    "xRGB": (1, 2, 3), "ARGB": (1, 2, 3), "xBGR": (3, 2, 1), "ABGR": (3, 2, 1),
    "RGB": (0, 1, 2), "BGR": (2, 1, 0), "AYUV": (1, 2, 3),
}
ALPHA_OFFSET = {"RGBx": 3, "RGBA": 3, "BGRx": 3, "BGRA": 3, "xRGB": 0, "ARGB": 0, "xBGR": 0, "ABGR": 0, "AYUV": 0}
PRESETS = {"none": 0, "heat": 1, "sepia": 2, "xray": 3, "xpro": 4, "yellowblue": 5}
PRESET_MAP_LUMA = {"heat": 1, "sepia": 1, "xray": 1, "xpro": 0, "yellowblue": 0}
OFF_EDGE = {"ignore": 0, "clamp": 1, "wrap": 2}


def build_port():
    subprocess.check_call(["make", "-s", "-C", HERE])
    return PORT_SO


def build_ref():
    import importlib.util
    spec = importlib.util.spec_from_file_location("b200vf_build_ref", os.path.join(HERE, "build_ref.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.build()


def have_ref():
    return os.path.exists(REF_SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def round_up_4(n):
    return (n + 3) & ~3


def coloreffects_tables():
    """The five 256x3 preset tables as {name: uint8[768]} from the committed
    golden fixture (generated from the reference by tests/golden/make_golden.py)."""
    path = os.path.join(HERE, "..", "tests", "golden", "coloreffects_tables.npz")
    z = np.load(path)
    return {k: z[k] for k in z.files}


class _Base:
    kind = None

    # -- helpers shared by both back-ends ---------------------------------
    def gaussblur(self, frame, sigma, p0=1):
        """frame: uint8 [h, stride] packed 4 B/px (stride >= 4*w given by shape[1]),
        returns the blurred copy. Emulates the element: copy in->out, then blur
        from/to COMP_DATA = plane + p0 with a zeroed slack after the frame (D5)."""
        raise NotImplementedError


class Port(_Base):
    kind = "port"

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build_port()
        self.lib = C.CDLL(PORT_SO)
        self.lib.oracle_gauss_kernel.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
        self.lib.oracle_rgb_to_hue.restype = C.c_int

    def bayer2rgb(self, src, width, height, fmt="bggr", out="RGBA"):
        src = _u8(src)
        stride = src.shape[1]
        dst = np.zeros((height, width * 4), np.uint8)
        r, g, b = RGB_OFFSETS[out]
        rc = self.lib.oracle_bayer2rgb(_p(dst), C.c_int(width * 4), _p(src), C.c_int(stride), width, height,
                                       BAYER_FORMATS[fmt], r, g, b)
        if rc != 0:
            raise ValueError("bayer2rgb: outside the reference's domain (even w>=4, h>=3)")
        return dst

    def rgb2bayer(self, src, width, height, fmt="bggr"):
        src = _u8(src)
        dst = np.zeros((height, round_up_4(width)), np.uint8)
        self.lib.oracle_rgb2bayer(_p(dst), C.c_int(dst.shape[1]), _p(src), C.c_int(src.shape[1]), width, height,
                                  BAYER_FORMATS[fmt])
        return dst

    def gauss_kernel(self, sigma):
        k = np.zeros(128, np.float32)
        s = np.zeros(128, np.float32)
        ws = self.lib.oracle_gauss_kernel(C.c_float(sigma), _p(k), _p(s))
        return k[:ws].copy(), s[:ws].copy()

    def gaussblur(self, frame, width, height, sigma, p0=1):
        frame = _u8(frame)
        stride = frame.shape[1]
        buf_in = np.zeros(frame.size + 64, np.uint8)
        buf_in[:frame.size] = frame.reshape(-1)
        buf_out = buf_in.copy()
        if np.float32(sigma) != 0.0:
            k, s = self.gauss_kernel(sigma)
            self.lib.oracle_gaussblur(C.c_void_p(buf_in.ctypes.data + p0), C.c_void_p(buf_out.ctypes.data + p0),
                                      width, height, stride, _p(k), _p(s), len(k))
        return buf_out[:frame.size].reshape(frame.shape)

    def _pt(self, fn, src, *args):
        src = np.ascontiguousarray(src, dtype=np.uint32)
        dst = np.zeros_like(src)
        fn(_p(src), _p(dst), src.size, *args)
        return dst

    def burn(self, src, adjustment=175):
        src = np.ascontiguousarray(src, dtype=np.uint32)
        dst = np.zeros_like(src)
        self.lib.oracle_burn(_p(dst), _p(src), int(adjustment), src.size)
        return dst

    def dodge(self, src):
        return self._pt(self.lib.oracle_dodge, src)

    def chromium(self, src, edge_a=200, edge_b=1):
        return self._pt(self.lib.oracle_chromium, src, int(edge_a), int(edge_b))

    def exclusion(self, src, factor=175):
        return self._pt(self.lib.oracle_exclusion, src, int(factor))

    def solarize(self, src, threshold=127, start=50, end=185):
        return self._pt(self.lib.oracle_solarize, src, int(threshold), int(start), int(end))

    def dilate(self, src, erode=False):
        src = np.ascontiguousarray(src, dtype=np.uint32)
        h, w = src.shape
        dst = np.zeros_like(src)
        self.lib.oracle_dilate(_p(src), _p(dst), w, h, int(bool(erode)))
        return dst

    def coloreffects(self, frame, width, height, fmt, preset):
        frame = _u8(frame).copy()
        if preset == "none":
            return frame
        table = coloreffects_tables()[preset]
        o = RGB_OFFSETS[fmt]
        ps = 3 if fmt in ("RGB", "BGR") else 4
        self.lib.oracle_coloreffects(_p(frame), width, height, frame.shape[1], ps, o[0], o[1], o[2],
                                     _p(table), PRESET_MAP_LUMA[preset], int(fmt == "AYUV"))
        return frame

    def chromahold(self, frame, width, height, fmt, target=(255, 0, 0), tolerance=30):
        frame = _u8(frame).copy()
        o = RGB_OFFSETS[fmt]
        self.lib.oracle_chromahold(_p(frame), width, height, frame.shape[1], o[0], o[1], o[2],
                                   int(target[0]), int(target[1]), int(target[2]), int(tolerance))
        return frame

    def rgb_to_hue(self, r, g, b):
        return self.lib.oracle_rgb_to_hue(int(r), int(g), int(b))

    def gt_map(self, element, width, height, props=None):
        props = props or {}
        m = np.zeros((height, width, 2), np.float64)
        fn = getattr(self.lib, "oracle_map_" + element, None)
        if fn is None:
            raise NotImplementedError(element)
        if props:
            raise NotImplementedError("properties for %s" % element)
        fn(_p(m), width, height)
        return m

    def remap(self, frame, gmap, width, height, pixel_stride, off_edge="ignore", is_ayuv=False):
        frame = _u8(frame)
        out = np.zeros_like(frame)
        gmap = np.ascontiguousarray(gmap, np.float64)
        self.lib.oracle_remap(_p(frame), _p(out), C.c_size_t(out.size), _p(gmap), width, height, pixel_stride,
                              frame.shape[1], OFF_EDGE[off_edge], int(is_ayuv))
        return out

    # ---- videofilters ------------------------------------
    def zebrastripe(self, frame, width, height, threshold=90, t=0, pixel_stride=1, luma_offset=0):
        """frame: uint8 [rows, stride] holding plane 0 (or the packed frame); luma sample of pixel i of row j =
        byte luma_offset + i*pixel_stride of row j. Returns the striped copy."""
        out = _u8(frame).copy()
        self.lib.oracle_zebrastripe(C.c_void_p(out.ctypes.data + luma_offset), pixel_stride, out.shape[1], width, height,
                                    self.lib.oracle_zebrastripe_y_threshold(threshold), t)
        return out

    def videodiff_luma(self, old, new, width, height, threshold=10, t=0):
        old, new = _u8(old), _u8(new)
        out = new.copy()
        self.lib.oracle_videodiff_luma(_p(out), _p(new), _p(old), new.shape[1], width, height, threshold, t)
        return out

    def sad_u8(self, a, b, width, height):
        a, b = _u8(a), _u8(b)
        self.lib.oracle_sad_u8.restype = C.c_uint32
        return int(self.lib.oracle_sad_u8(_p(a), a.shape[1], _p(b), b.shape[1], width, height))

    def videoanalyse(self, luma, width, height):
        """(luma-average, luma-variance) as the element posts them"""
        luma = _u8(luma)
        a, v = C.c_double(0), C.c_double(0)
        self.lib.oracle_videoanalyse(_p(luma), luma.shape[1], width, height, C.byref(a), C.byref(v))
        return a.value, v.value

    # simplevideomark / simplevideomarkdetect: a pure-Python restatement (a few hundred samples per frame) of
    # gst_video_mark_yuv (gst/videosignal/gstsimplevideomark.c:348-462) and gst_video_detect_yuv
    # (gstsimplevideomarkdetect.c:420-565): the same walk over the boxes, statement for statement.
    @staticmethod
    def _calculate_pw(pw, x, width):                          # :336-345
        if x < 0:
            return pw + x
        if x + pw > width:
            return width - x
        return pw

    def videomark(self, frame, pixel_stride, width, height, pw=4, ph=16, pc=4, pdc=5, data=10, left=0, bottom=0):
        out = np.ascontiguousarray(frame, np.uint8).copy()
        flat = out.reshape(-1)
        rs = out.shape[1]
        offset = rs * (height - ph - bottom) + pixel_stride * left
        x, y, total = left, height - ph - bottom, pc + pdc
        if (x + pw * total) < 0 or x > width or (y + height) < 0 or y > height:
            return out
        offset = max(offset, 0)
        if y < 0:
            ph += y
        elif y + ph > height:
            ph = height - y
        if ph < 0:
            return out
        d = offset

        def box(d, w, color):
            for i in range(ph):
                for j in range(w):
                    flat[d + i * rs + pixel_stride * j] = color
        for i in range(pc):
            dpw = self._calculate_pw(pw, x, width)
            if dpw < 0:
                continue
            box(d, dpw, 255 if i & 1 else 0)
            d += pixel_stride * dpw
            x += dpw
            if (x + pw * (total - i - 1)) < 0 or x >= width:
                return out
        shift = 1 << (pdc - 1) if pdc > 0 else 0
        for i in range(pdc):
            dpw = self._calculate_pw(pw, x, width)
            if dpw < 0:
                continue
            box(d, dpw, 255 if data & shift else 0)
            shift >>= 1
            d += pixel_stride * dpw
            x += dpw
            if (x + pw * (pdc - i - 1)) < 0 or x >= width:
                return out
        return out

    def videomarkdetect(self, frame, pixel_stride, width, height, pw=4, ph=16, pc=4, pdc=5, center=0.5, sensitivity=0.3, left=0,
                        bottom=0, in_pattern=False):
        f = np.ascontiguousarray(frame, np.uint8)
        flat = f.reshape(-1)
        rs = f.shape[1]
        offset = rs * (height - ph - bottom) + pixel_stride * left
        x, y, total = left, height - ph - bottom, pc + pdc
        if (x + pw * total) < 0 or x > width or (y + height) < 0 or y > height:
            return in_pattern, False, 0
        offset = max(offset, 0)
        if y < 0:
            ph += y
        elif y + ph > height:
            ph = height - y
        if ph < 0:
            return in_pattern, False, 0
        d = offset

        def brightness(d):                                    # :392-407; samples past the plane count as 0 here (undefined there)
            s = 0
            for i in range(ph):
                for j in range(pw):
                    o = d + i * rs + pixel_stride * j
                    s += int(flat[o]) if o < flat.size else 0
            den = 255.0 * pw * ph
            return s / den if den else float("nan")
        for i in range(pc):
            b = brightness(d)
            if i & 1:
                if b < center + sensitivity:
                    return False, bool(in_pattern), 0
            elif b > center - sensitivity:
                return False, bool(in_pattern), 0
            dpw = self._calculate_pw(pw, x, width)
            if dpw < 0:
                continue
            d += pixel_stride * dpw
            x += dpw
            if (x + pw * (total - i - 1)) < 0 or x >= width:
                break
        data = 0
        for i in range(pdc):
            b = brightness(d)
            data <<= 1
            if b > center:
                data |= 1
            dpw = self._calculate_pw(pw, x, width)
            if dpw < 0:
                continue
            d += pixel_stride * dpw
            x += dpw
            if (x + pw * (pdc - i - 1)) < 0 or x >= width:
                break
        return True, True, data

    def smooth_plane(self, plane, width, height, tolerance=8, filtersize=3, prefill=0):
        """returns the filtered plane; rows the reference never writes keep `prefill`"""
        plane = _u8(plane)
        out = np.full_like(plane, prefill)
        self.lib.oracle_smooth_plane(_p(out), _p(plane), width, height, plane.shape[1], plane.shape[1], tolerance, filtersize)
        return out

    def scenechange_run(self, scores):
        """the element's decision over a sequence of frame scores -> list of bools"""
        class S(C.Structure):
            _fields_ = [("n_diffs", C.c_int), ("diffs", C.c_double * 5)]
        st = S()
        self.lib.oracle_scenechange_update.argtypes = [C.c_void_p, C.c_double]
        return [bool(self.lib.oracle_scenechange_update(C.byref(st), float(x))) for x in scores]


class _GT(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("pixel_stride", C.c_int), ("row_stride", C.c_int),
                ("precalc_map", C.c_int), ("needs_remap", C.c_int), ("off_edge_pixels", C.c_int),
                ("map", C.c_void_p)]


class Ref(_Base):
    kind = "reference"

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " (run oracle/build_ref.py where /root/reference is mounted)")
        self.lib = C.CDLL(REF_SO)
        self.lib.ref_gauss_kernel.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
        self.lib.ref_gaussblur.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float]
        self.lib.ref_rgb_to_hue.restype = C.c_int

    def bayer2rgb(self, src, width, height, fmt="bggr", out="RGBA"):
        src = _u8(src)
        dst = np.zeros((height, width * 4), np.uint8)
        r, g, b = RGB_OFFSETS[out]
        self.lib.ref_bayer2rgb(_p(dst), C.c_int(width * 4), _p(src), C.c_int(src.shape[1]), width, height,
                               BAYER_FORMATS[fmt], r, g, b)
        return dst

    def rgb2bayer(self, src, width, height, fmt="bggr"):
        src = _u8(src)
        dst = np.zeros((height, round_up_4(width)), np.uint8)
        self.lib.ref_rgb2bayer(_p(dst), _p(src), C.c_int(src.shape[1]), width, height, BAYER_FORMATS[fmt])
        return dst

    def gauss_kernel(self, sigma):
        k = np.zeros(128, np.float32)
        s = np.zeros(128, np.float32)
        ws = self.lib.ref_gauss_kernel(C.c_float(sigma), _p(k), _p(s))
        return k[:ws].copy(), s[:ws].copy()

    def gaussblur(self, frame, width, height, sigma, p0=1):
        frame = _u8(frame)
        buf_in = np.zeros(frame.size + 64, np.uint8)
        buf_in[:frame.size] = frame.reshape(-1)
        buf_out = buf_in.copy()
        self.lib.ref_gaussblur(C.c_void_p(buf_in.ctypes.data + p0), C.c_void_p(buf_out.ctypes.data + p0),
                               width, height, frame.shape[1], C.c_float(sigma))
        return buf_out[:frame.size].reshape(frame.shape)

    def _pt(self, fn, src, *args):
        src = np.ascontiguousarray(src, dtype=np.uint32)
        dst = np.zeros_like(src)
        fn(_p(src), _p(dst), src.size, *args)
        return dst

    def burn(self, src, adjustment=175):
        src = np.ascontiguousarray(src, dtype=np.uint32)
        dst = np.zeros_like(src)
        self.lib.ref_burn(_p(dst), _p(src), int(adjustment), src.size)
        return dst

    def dodge(self, src):
        return self._pt(self.lib.ref_dodge, src)

    def chromium(self, src, edge_a=200, edge_b=1):
        return self._pt(self.lib.ref_chromium, src, int(edge_a), int(edge_b))

    def exclusion(self, src, factor=175):
        return self._pt(self.lib.ref_exclusion, src, int(factor))

    def solarize(self, src, threshold=127, start=50, end=185):
        return self._pt(self.lib.ref_solarize, src, int(threshold), int(start), int(end))

    def dilate(self, src, erode=False):
        src = np.ascontiguousarray(src, dtype=np.uint32)
        h, w = src.shape
        dst = np.zeros_like(src)
        self.lib.ref_dilate(_p(src), _p(dst), src.size, w, h, int(bool(erode)))
        return dst

    def coloreffects_table(self, preset):
        t = np.zeros(768, np.uint8)
        ml = C.c_int(0)
        ok = self.lib.ref_coloreffects_table(PRESETS[preset], _p(t), C.byref(ml))
        return (t, ml.value) if ok else (None, 0)

    def coloreffects(self, frame, width, height, fmt, preset):
        frame = _u8(frame).copy()
        o = RGB_OFFSETS[fmt]
        ps = 3 if fmt in ("RGB", "BGR") else 4
        self.lib.ref_coloreffects(_p(frame), width, height, frame.shape[1], ps, o[0], o[1], o[2],
                                  PRESETS[preset], int(fmt == "AYUV"))
        return frame

    def chromahold(self, frame, width, height, fmt, target=(255, 0, 0), tolerance=30):
        frame = _u8(frame).copy()
        o = RGB_OFFSETS[fmt]
        self.lib.ref_chromahold(_p(frame), width, height, frame.shape[1], o[0], o[1], o[2], ALPHA_OFFSET[fmt],
                                int(target[0]), int(target[1]), int(target[2]), int(tolerance))
        return frame

    def rgb_to_hue(self, r, g, b):
        return self.lib.ref_rgb_to_hue(int(r), int(g), int(b))

    # geometrictransform: element instance = calloc'd blob of the element's struct
    def _gt_elem(self, element, width, height, pixel_stride=4, row_stride=None, off_edge="ignore", props=None):
        props = props or {}
        size_fn = getattr(self.lib, "ref_gt_%s_size" % element)
        size_fn.restype = C.c_size_t
        blob = C.create_string_buffer(size_fn())
        gt = C.cast(blob, C.POINTER(_GT)).contents
        gt.width, gt.height = width, height
        gt.pixel_stride = pixel_stride
        gt.row_stride = row_stride if row_stride is not None else width * pixel_stride
        gt.off_edge_pixels = OFF_EDGE[off_edge]
        for k, v in props.items():
            f = getattr(self.lib, "ref_gt_%s_set_%s" % (element, k))
            f.argtypes = [C.c_void_p, C.c_double]
            f(blob, float(v))
        getattr(self.lib, "ref_gt_%s_prepare" % element)(blob)
        return blob, gt

    def gt_map(self, element, width, height, props=None):
        """props: the element's struct fields (x_center, zoom, ...). NOTE: the reference sets its
        defaults in each element's init(); this harness starts from a zeroed struct, so callers
        pass every field (tests/refprops.py holds the defaults)."""
        blob, gt = self._gt_elem(element, width, height, props=props)
        mf = getattr(self.lib, "ref_gt_%s_map" % element)
        mf.restype = C.c_void_p
        m = np.zeros((height, width, 2), np.float64)
        ok = self.lib.ref_gt_generate_map(C.c_void_p(mf()), blob, _p(m))
        if not ok:
            raise RuntimeError("map_func failed")
        return m

    def remap(self, frame, gmap, width, height, pixel_stride, off_edge="ignore", is_ayuv=False):
        frame = _u8(frame)
        out = np.zeros_like(frame)
        gt = _GT(width, height, pixel_stride, frame.shape[1], 1, 0, OFF_EDGE[off_edge], None)
        gmap = np.ascontiguousarray(gmap, np.float64)
        self.lib.ref_gt_apply_map(C.byref(gt), _p(gmap), _p(frame), _p(out), C.c_size_t(out.size), int(is_ayuv))
        return out

    # ---- videofilters ------------------------------------
    def zebrastripe(self, frame, width, height, threshold=90, t=0, pixel_stride=1, luma_offset=0):
        """the reference's loop works from frame->data[0] with (offset, pixel_stride, y_position): UYVY has
        offset 1, AYUV y_position 1 (gstzebrastripe.c:232-237); both are `luma_offset` here."""
        out = _u8(frame).copy()
        off, ypos = (0, luma_offset) if pixel_stride == 4 else (luma_offset, 0)
        self.lib.ref_zebrastripe(_p(out), out.shape[1], width, height, self.lib.ref_zebrastripe_y_threshold(threshold), t,
                                 off, pixel_stride, ypos)
        return out

    def videodiff_luma(self, old, new, width, height, threshold=10, t=0):
        """through the reference's three-plane function with 1x1 dummy chroma planes"""
        old, new = _u8(old), _u8(new)
        out = np.zeros_like(new)
        dummy = [np.zeros((1, 4), np.uint8) for _ in range(6)]
        P = C.c_void_p * 3
        outs = P(out.ctypes.data, dummy[0].ctypes.data, dummy[1].ctypes.data)
        ins = P(new.ctypes.data, dummy[2].ctypes.data, dummy[3].ctypes.data)
        olds = P(old.ctypes.data, dummy[4].ctypes.data, dummy[5].ctypes.data)
        strides = (C.c_int * 3)(new.shape[1], 4, 4)
        self.lib.ref_videodiff(outs, ins, olds, strides, width, height, 1, 1, threshold, t)
        return out

    def sad_u8(self, a, b, width, height):
        a, b = _u8(a), _u8(b)
        sad = C.c_uint32(0)
        self.lib.ref_scenechange_score.restype = C.c_double
        self.lib.ref_scenechange_score(_p(a), a.shape[1], _p(b), b.shape[1], width, height, C.byref(sad))
        return int(sad.value)

    def videomark(self, frame, pixel_stride, width, height, pw=4, ph=16, pc=4, pdc=5, data=10, left=0, bottom=0):
        """gst_video_mark_yuv on a copy of `frame` (rows of the plane; luma at byte 0 of every pixel_stride bytes)"""
        out = np.ascontiguousarray(frame, np.uint8).copy()
        self.lib.ref_videomark(_p(out), out.shape[1], pixel_stride, width, height, pw, ph, pc, pdc, C.c_uint64(data), left, bottom)
        return out

    def videomarkdetect(self, frame, pixel_stride, width, height, pw=4, ph=16, pc=4, pdc=5, center=0.5, sensitivity=0.3, left=0,
                        bottom=0, in_pattern=False):
        """gst_video_detect_yuv on one frame -> (in_pattern, message posted, data)"""
        f = np.ascontiguousarray(frame, np.uint8)
        ip, data = C.c_int(int(in_pattern)), C.c_uint64(0)
        self.lib.ref_videomarkdetect.restype = C.c_int
        n = self.lib.ref_videomarkdetect(_p(f), f.shape[1], pixel_stride, width, height, pw, ph, pc, pdc, C.c_double(center),
                                         C.c_double(sensitivity), left, bottom, C.byref(ip), C.byref(data))
        return bool(ip.value), bool(n), int(data.value)

    def videoanalyse(self, luma, width, height):
        luma = _u8(luma)
        a, v = C.c_double(0), C.c_double(0)
        self.lib.ref_videoanalyse(_p(luma), luma.shape[1], width, height, C.byref(a), C.byref(v))
        return a.value, v.value

    def smooth_plane(self, plane, width, height, tolerance=8, filtersize=3, prefill=0):
        plane = _u8(plane)
        out = np.full_like(plane, prefill)
        self.lib.ref_smooth_plane(_p(out), _p(plane), width, height, plane.shape[1], plane.shape[1], tolerance, filtersize)
        return out

    def scenechange_run(self, scores):
        class S(C.Structure):
            _fields_ = [("n_diffs", C.c_int), ("diffs", C.c_double * 5)]
        st = S()
        self.lib.ref_scenechange_update.argtypes = [C.c_void_p, C.c_double]
        return [bool(self.lib.ref_scenechange_update(C.byref(st), float(x))) for x in scores]


_cache = {}


def get(kind="port"):
    if kind not in _cache:
        _cache[kind] = Port() if kind == "port" else Ref()
    return _cache[kind]


def best():
    """The reference build when present, else the port."""
    return get("reference") if have_ref() else get("port")
