"""ctypes binding of libb200vf.so (include/b200vf.h).

Thin by design: every call below is one C-ABI call; the product is the shared
library.  Import fails loudly when the library has not been built, and
`Context()` fails loudly when there is no sm_100 GPU - there is no CPU path.
"""
import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "..", "lib", "libb200vf.so")


class B200vfError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("b200vf status %d: %s" % (status, message))
        self.status = status


OK, E_INVAL, E_NO_DEVICE, E_CUDA, E_NOMEM, E_UNSUPPORTED, E_NOT_NEGOTIATED, E_NCCL, E_PROPERTY = 0, -1, -2, -3, -4, -5, -6, -7, -8

if not os.path.exists(LIB_PATH):
    raise ImportError("libb200vf.so is not built (%s); run `make -C gst-plugins-bad_b200/csrc` or "
                      "__graft_entry__.build() - the CUDA extension is mandatory, there is no fallback" % LIB_PATH)
lib = C.CDLL(os.path.abspath(LIB_PATH))

_vp, _i, _sz, _u32 = C.c_void_p, C.c_int, C.c_size_t, C.c_uint32
_LUT = C.c_uint8 * 1024

_SIGS = {
    "b200vf_version": (_i, []),
    "b200vf_last_error": (C.c_char_p, []),
    "b200vf_status_string": (C.c_char_p, [_i]),
    "b200vf_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "b200vf_ctx_destroy": (None, [_vp]),
    "b200vf_ctx_device": (_i, [_vp]),
    "b200vf_ctx_stream": (_vp, [_vp]),
    "b200vf_ctx_sm_count": (_i, [_vp]),
    "b200vf_ctx_synchronize": (_i, [_vp, _vp]),
    "b200vf_ctx_launch_count": (C.c_uint64, [_vp]),
    "b200vf_ctx_last_kernel": (C.c_char_p, [_vp]),
    "b200vf_ctx_set_variant": (_i, [_vp, _i]),
    "b200vf_pool_create": (_i, [_vp, _sz, _i, C.POINTER(_vp)]),
    "b200vf_pool_destroy": (None, [_vp]),
    "b200vf_pool_acquire": (_i, [_vp, C.POINTER(_i)]),
    "b200vf_pool_release": (_i, [_vp, _i]),
    "b200vf_pool_device_ptr": (_vp, [_vp, _i]),
    "b200vf_pool_host_ptr": (_vp, [_vp, _i]),
    "b200vf_pool_buf_bytes": (_sz, [_vp]),
    "b200vf_pool_buf_pitch": (_sz, [_vp]),
    "b200vf_pool_upload": (_i, [_vp, _i, _vp, _sz, _vp]),
    "b200vf_pool_download": (_i, [_vp, _i, _vp, _sz, _vp]),
    "b200vf_memory_new": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "b200vf_pool_acquire_memory": (_i, [_vp, C.POINTER(_vp)]),
    "b200vf_memory_ref": (_vp, [_vp]),
    "b200vf_memory_unref": (None, [_vp]),
    "b200vf_memory_size": (_sz, [_vp]),
    "b200vf_memory_flags": (C.c_uint, [_vp]),
    "b200vf_memory_is_writable": (_i, [_vp]),
    "b200vf_memory_pending_stages": (_i, [_vp]),
    "b200vf_memory_map": (_i, [_vp, _i, C.POINTER(_vp), _vp]),
    "b200vf_memory_unmap": (_i, [_vp]),
    "b200vf_ctx_transfer_counts": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "b200vf_element_transform": (_i, [_vp, _vp, _vp, _i, _vp]),
    "b200vf_element_set_host_mode": (_i, [_vp, _i]),
    "b200vf_host_pin_cache_clear": (_i, []),
    "b200vf_element_default_layout": (_i, [_vp, _i, _vp]),
    "b200vf_element_transform_host_layout": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "b200vf_malloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "b200vf_free": (_i, [_vp, _vp]),
    "b200vf_host_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "b200vf_host_free": (_i, [_vp]),
    "b200vf_memcpy_h2d": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "b200vf_memcpy_d2h": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "b200vf_bayer2rgb": (_i, [_vp, _vp, _i, _sz, _vp, _i, _sz, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "b200vf_bayer2rgb_shard": (_i, [_vp, _vp, _i, _sz, _vp, _i, _sz, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "b200vf_bayer2rgb_shard_fused": (_i, [_vp, _vp, _i, _sz, _vp, _i, _sz, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "b200vf_rgb2bayer": (_i, [_vp, _vp, _i, _sz, _vp, _i, _sz, _i, _i, _i, _i, _vp]),
    "b200vf_lut4": (_i, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "b200vf_lut_burn": (_i, [_i, _vp]),
    "b200vf_lut_dodge": (_i, [_vp]),
    "b200vf_lut_chromium": (_i, [_i, _i, _vp]),
    "b200vf_lut_solarize": (_i, [_i, _i, _i, _vp]),
    "b200vf_lut_compose": (_i, [_vp, _vp, _vp]),
    "b200vf_exclusion": (_i, [_vp, _vp, _vp, _sz, _i, _vp]),
    "b200vf_dilate": (_i, [_vp, _vp, _vp, _i, _i, _sz, _i, _i, _vp, _vp]),
    "b200vf_gauss_kernel": (_i, [C.c_float, _vp, _vp, _i]),
    "b200vf_gaussblur_halo_rows": (_i, [_i, _i, _i, _i]),
    "b200vf_gaussblur": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _sz, _i, _i, _vp, _vp, _i, _i, _vp]),
    "b200vf_gauss_selftest_div": (_i, [_vp, C.c_float, C.c_uint32, C.c_uint32, _vp]),
    "b200vf_gauss_selftest_finish": (_i, [_vp, C.c_uint32, C.c_uint32, _vp]),
    "b200vf_gauss_selftest_div1": (_i, [_vp, C.c_float, C.c_float, C.c_uint32, C.c_uint32, _vp]),
    "b200vf_gauss_div1_constant": (_i, [C.c_float, _vp]),
    "b200vf_coloreffects_table": (_i, [_i, C.POINTER(_vp), C.POINTER(_i)]),
    "b200vf_coloreffects_rgb": (_i, [_vp, _vp, _i, _i, _i, _sz, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "b200vf_coloreffects_ayuv": (_i, [_vp, _vp, _i, _i, _i, _sz, _i, _i, _i, _i, _vp, _i, _vp]),
    "b200vf_chromahold": (_i, [_vp, _vp, _i, _i, _i, _sz, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "b200vf_gt_build_map": (_i, [C.c_char_p, _i, _i, C.POINTER(C.c_char_p), C.POINTER(C.c_double), _i, _vp]),
    "b200vf_gt_device_map_supported": (_i, [C.c_char_p]),
    "b200vf_gt_device_last_uncertain": (C.c_longlong, []),
    "b200vf_gt_build_index_device": (_i, [_vp, C.c_char_p, _i, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "b200vf_gt_resolve_map": (_i, [_vp, _i, _i, _i, _vp]),
    "b200vf_remap": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _sz, _i, _u32, _vp]),
    "b200vf_diffuse_draw": (None, [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(_i), C.POINTER(C.c_double)]),
    "b200vf_diffuse_tables": (_i, [C.c_double, _vp, _vp]),
    "b200vf_diffuse": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _sz, _sz, _i, _vp, _vp, _i, _u32, C.c_uint64, C.c_uint64, _vp]),
    "b200vf_element_set_rng_seed": (_i, [_vp, C.c_uint64, C.c_uint64]),
    "b200vf_element_get_rng_state": (_i, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "b200vf_gt_packed_bound": (_sz, [_i, _i]),
    "b200vf_gt_pack_index": (_i, [_vp, _i, _i, _vp, _sz, C.POINTER(_sz), C.POINTER(_sz)]),
    "b200vf_gt_unpack_index": (_i, [_vp, _sz, _i, _i, _vp]),
    "b200vf_remap_packed": (_i, [_vp, _vp, _vp, _vp, _i, _i, _sz, _i, _u32, _vp]),
    "b200vf_zebrastripe_y_threshold": (_i, [_i]),
    "b200vf_zebrastripe": (_i, [_vp, _vp, _i, _i, _sz, _i, _i, _i, _i, _i, _vp]),
    "b200vf_videodiff_luma": (_i, [_vp, _vp, _i, _sz, _vp, _i, _sz, _vp, _i, _sz, _i, _i, _i, _i, _i, _vp]),
    "b200vf_sad_u8": (_i, [_vp, _vp, _i, _sz, _vp, _i, _sz, _i, _i, _i, _vp, _vp]),
    "b200vf_luma_moments": (_i, [_vp, _vp, _i, _sz, _i, _i, _i, _vp, _vp]),
    "b200vf_videoanalyse_finish": (_i, [C.c_uint64, C.c_uint64, _i, _i, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "b200vf_videomark_draw": (_i, [_vp, _vp, _i, _i, _sz, _i, _i, _i, _vp, C.c_uint64, _vp]),
    "b200vf_videomark_box_sums": (_i, [_vp, _vp, _i, _i, _sz, _i, _i, _i, _vp, _vp, C.POINTER(_i), _vp]),
    "b200vf_videomark_detect_decide": (_i, [_vp, _i, _i, _i, _i, _vp, C.c_double, C.c_double, C.POINTER(_i), C.POINTER(_i), C.POINTER(C.c_uint64)]),
    "b200vf_smooth_plane": (_i, [_vp, _vp, _i, _sz, _vp, _i, _sz, _i, _i, _i, _i, _i, _vp]),
    "b200vf_scenechange_reset": (_i, [_vp]),
    "b200vf_scenechange_update": (_i, [_vp, C.c_double, C.POINTER(_i)]),
    "b200vf_bayer2rgb_fused": (_i, [_vp, _vp, _i, _sz, _vp, _i, _sz, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "b200vf_comm_unique_id": (_i, [_vp]),
    "b200vf_comm_create": (_i, [_vp, _vp, _i, _i, C.POINTER(_vp)]),
    "b200vf_comm_destroy": (None, [_vp]),
    "b200vf_shard_rows": (_i, [_i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "b200vf_comm_halo_exchange": (_i, [_vp, _vp, _sz, _i, _i, _sz, _i, _vp]),
    "b200vf_comm_halo_begin": (_i, [_vp, _vp, _sz, _i, _i, _sz, _i, _vp]),
    "b200vf_comm_halo_end": (_i, [_vp, _vp]),
    "b200vf_comm_barrier": (_i, [_vp, _vp]),
    "b200vf_comm_allgather_rows": (_i, [_vp, _vp, _sz, _i, _sz, _i, _vp]),
    "b200vf_comm_exchange_rows": (_i, [_vp, _vp, _sz, _i, _vp, _vp, _sz, _i, _vp]),
    "b200vf_gt_index_row_range": (_i, [_vp, _sz, _i, C.POINTER(_i), C.POINTER(_i)]),
    "b200vf_factory_count": (_i, []),
    "b200vf_factory_get": (_i, [_i, _vp]),
    "b200vf_factory_find": (_i, [C.c_char_p, _vp]),
    "b200vf_factory_property": (_i, [C.c_char_p, _i, _vp]),
    "b200vf_factory_format": (C.c_char_p, [C.c_char_p, _i]),
    "b200vf_element_factory_make": (_i, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "b200vf_element_destroy": (None, [_vp]),
    "b200vf_element_factory_name": (C.c_char_p, [_vp]),
    "b200vf_element_set_property": (_i, [_vp, C.c_char_p, C.c_double]),
    "b200vf_element_set_property_string": (_i, [_vp, C.c_char_p, C.c_char_p]),
    "b200vf_element_get_property": (_i, [_vp, C.c_char_p, C.POINTER(C.c_double)]),
    "b200vf_element_set_caps": (_i, [_vp, C.c_char_p, C.c_char_p, _i, _i]),
    "b200vf_element_unit_size": (_i, [_vp, C.POINTER(_sz), C.POINTER(_sz)]),
    "b200vf_element_transform_host": (_i, [_vp, _vp, _vp, _i]),
    "b200vf_element_transform_device": (_i, [_vp, _vp, _vp, _i, _vp]),
    "b200vf_element_last_values": (_i, [_vp, _vp, _i]),
    "b200vf_element_last_events": (_i, [_vp, _vp, _i]),
}

MISSING = []
for _name, (_res, _args) in _SIGS.items():
    try:
        _f = getattr(lib, _name)
    except AttributeError:
        MISSING.append(_name)
        continue
    _f.restype = _res
    _f.argtypes = _args


def declared_symbols():
    return sorted(_SIGS)


def check(status):
    if status != OK:
        raise B200vfError(status, lib.b200vf_last_error().decode("utf-8", "replace"))


def _ptr(x):
    """Device pointer from a DeviceBuffer, an int, or anything with data_ptr() (torch tensors)."""
    if x is None:
        return None
    if isinstance(x, DeviceBuffer):
        return x.ptr
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return int(x)


def _hptr(a):
    return a.ctypes.data_as(_vp) if isinstance(a, np.ndarray) else a


def _lut_arg(lut):
    if lut is None:
        return None
    a = np.ascontiguousarray(lut, dtype=np.uint8)
    assert a.shape == (4, 256), a.shape
    return a


class Context:
    def __init__(self, device=0):
        h = _vp()
        check(lib.b200vf_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device
        self._children = weakref.WeakSet()      # elements / buffers / comms: closed before the context

    def close(self):
        if self.h:
            for c in list(self._children):
                try:
                    c.close() if hasattr(c, "close") else c.free()
                except Exception:
                    pass
            lib.b200vf_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return lib.b200vf_ctx_stream(self.h)

    @property
    def sm_count(self):
        return lib.b200vf_ctx_sm_count(self.h)

    def synchronize(self, stream=None):
        check(lib.b200vf_ctx_synchronize(self.h, stream))

    def launch_count(self):
        return int(lib.b200vf_ctx_launch_count(self.h))

    def last_kernel(self):
        return lib.b200vf_ctx_last_kernel(self.h).decode()

    def transfer_counts(self):
        """(h2d_count, h2d_bytes, d2h_count, d2h_bytes) of the copies made through Memory objects so far"""
        v = [C.c_uint64(0) for _ in range(4)]
        check(lib.b200vf_ctx_transfer_counts(self.h, *[C.byref(x) for x in v]))
        return tuple(int(x.value) for x in v)

    def memory(self, nbytes):
        return Memory(self, nbytes)

    def pool(self, buf_bytes, n_bufs):
        return Pool(self, buf_bytes, n_bufs)

    def set_variant(self, v):
        check(lib.b200vf_ctx_set_variant(self.h, {"auto": 0, "direct": 1, "tma": 2}.get(v, v)))

    # -- memory -----------------------------------------------------------
    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def upload(self, array, stream=None):
        a = np.ascontiguousarray(array)
        b = DeviceBuffer(self, a.nbytes)
        check(lib.b200vf_memcpy_h2d(self.h, b.ptr, _hptr(a), a.nbytes, stream))
        self.synchronize(stream)
        return b

    def download(self, buf, nbytes=None, dtype=np.uint8, stream=None, offset=0):
        n = nbytes if nbytes is not None else buf.nbytes - offset
        out = np.empty(n, np.uint8)
        check(lib.b200vf_memcpy_d2h(self.h, _hptr(out), _ptr(buf) + offset, n, stream))
        self.synchronize(stream)
        return out.view(dtype)

    # -- ops (device pointers) ----------------------------------------------
    def bayer2rgb(self, src, src_stride, dst, dst_stride, width, height, pattern, offs, nframes=1,
                  src_frame_stride=None, dst_frame_stride=None, stream=None):
        sfs = src_frame_stride if src_frame_stride is not None else src_stride * height
        dfs = dst_frame_stride if dst_frame_stride is not None else dst_stride * height
        check(lib.b200vf_bayer2rgb(self.h, _ptr(src), src_stride, sfs, _ptr(dst), dst_stride, dfs, width, height,
                                   nframes, pattern, offs[0], offs[1], offs[2], stream))

    def bayer2rgb_shard(self, src, src_stride, dst, dst_stride, width, full_height, row0, rows, pattern, offs,
                        nframes=1, src_frame_stride=0, dst_frame_stride=0, stream=None):
        check(lib.b200vf_bayer2rgb_shard(self.h, _ptr(src), src_stride, src_frame_stride, _ptr(dst), dst_stride,
                                         dst_frame_stride, width, full_height, row0, rows, nframes, pattern,
                                         offs[0], offs[1], offs[2], stream))

    def bayer2rgb_fused(self, src, src_stride, dst, dst_stride, width, height, pattern, offs, luma_table=None,
                        lut=None, nframes=1, src_frame_stride=None, dst_frame_stride=None, stream=None):
        sfs = src_frame_stride if src_frame_stride is not None else src_stride * height
        dfs = dst_frame_stride if dst_frame_stride is not None else dst_stride * height
        lt = None if luma_table is None else np.ascontiguousarray(luma_table, np.uint8)
        la = _lut_arg(lut)
        check(lib.b200vf_bayer2rgb_fused(self.h, _ptr(src), src_stride, sfs, _ptr(dst), dst_stride, dfs, width,
                                         height, nframes, pattern, offs[0], offs[1], offs[2],
                                         None if lt is None else _hptr(lt), None if la is None else _hptr(la), stream))

    def bayer2rgb_shard_fused(self, src, src_stride, dst, dst_stride, width, full_height, row0, rows, pattern, offs,
                              luma_table=None, lut=None, nframes=1, src_frame_stride=0, dst_frame_stride=0, stream=None):
        lt = None if luma_table is None else np.ascontiguousarray(luma_table, np.uint8)
        la = _lut_arg(lut)
        check(lib.b200vf_bayer2rgb_shard_fused(self.h, _ptr(src), src_stride, src_frame_stride, _ptr(dst), dst_stride,
                                               dst_frame_stride, width, full_height, row0, rows, nframes, pattern,
                                               offs[0], offs[1], offs[2], None if lt is None else _hptr(lt),
                                               None if la is None else _hptr(la), stream))

    def rgb2bayer(self, src, src_stride, dst, dst_stride, width, height, pattern, nframes=1, stream=None):
        check(lib.b200vf_rgb2bayer(self.h, _ptr(src), src_stride, src_stride * height, _ptr(dst), dst_stride,
                                   dst_stride * height, width, height, nframes, pattern, stream))

    def lut4(self, src, dst, npix, lut, stream=None):
        la = _lut_arg(lut)
        check(lib.b200vf_lut4(self.h, _ptr(src), _ptr(dst), npix, _hptr(la), stream))

    def exclusion(self, src, dst, npix, factor, stream=None):
        check(lib.b200vf_exclusion(self.h, _ptr(src), _ptr(dst), npix, factor, stream))

    def dilate(self, src, dst, width, height, erode=False, nframes=1, below=None, stream=None):
        check(lib.b200vf_dilate(self.h, _ptr(src), _ptr(dst), width, height, 4 * width * height, nframes,
                                int(bool(erode)), _ptr(below), stream))

    def gaussblur(self, src, dst, width, height, stride, p0, kernel, kernel_sum, exact=True, nframes=1,
                  frame_stride=None, row0=0, rows=None, full_height=None, stream=None):
        k = np.ascontiguousarray(kernel, np.float32)
        ks = np.ascontiguousarray(kernel_sum, np.float32)
        fh = full_height if full_height is not None else height
        rws = rows if rows is not None else height
        fs = frame_stride if frame_stride is not None else stride * rws
        check(lib.b200vf_gaussblur(self.h, _ptr(src), _ptr(dst), width, fh, row0, rws, stride, fs, nframes, p0,
                                   _hptr(k), _hptr(ks), len(k), int(bool(exact)), stream))

    def gauss_selftest_div(self, divisor, lo_bits, hi_bits):
        """mismatches between the blur's reciprocal-based division and IEEE a / divisor over fp32 bit patterns"""
        bad = C.c_ulonglong(0)
        check(lib.b200vf_gauss_selftest_div(self.h, float(divisor), int(lo_bits), int(hi_bits), C.byref(bad)))
        return bad.value

    # ---- videofiltersbad plugin
    def zebrastripe(self, luma, pixel_stride, row_stride, width, height, threshold=90, t=0, nframes=1, frame_stride=None,
                    stream=None):
        fs = frame_stride if frame_stride is not None else row_stride * height
        check(lib.b200vf_zebrastripe(self.h, _ptr(luma), pixel_stride, row_stride, fs, nframes, width, height,
                                     lib.b200vf_zebrastripe_y_threshold(threshold), t, stream))

    def videodiff_luma(self, old, new, out, stride, width, height, threshold=10, t=0, nframes=1, frame_stride=None, stream=None):
        fs = frame_stride if frame_stride is not None else stride * height
        check(lib.b200vf_videodiff_luma(self.h, _ptr(old), stride, fs, _ptr(new), stride, fs, _ptr(out), stride, fs,
                                        width, height, nframes, threshold, t, stream))

    def sad_u8(self, a, b, stride, width, height, sums, nframes=1, frame_stride=None, stream=None):
        fs = frame_stride if frame_stride is not None else stride * height
        check(lib.b200vf_sad_u8(self.h, _ptr(a), stride, fs, _ptr(b), stride, fs, width, height, nframes, _ptr(sums), stream))

    def luma_moments(self, luma, stride, width, height, sums, nframes=1, frame_stride=None, stream=None):
        fs = frame_stride if frame_stride is not None else stride * height
        check(lib.b200vf_luma_moments(self.h, _ptr(luma), stride, fs, width, height, nframes, _ptr(sums), stream))

    def videomark_draw(self, luma, pixel_stride, row_stride, width, height, params, pattern_data=10, nframes=1, frame_stride=None, stream=None):
        fs = frame_stride if frame_stride is not None else row_stride * height
        check(lib.b200vf_videomark_draw(self.h, _ptr(luma), pixel_stride, row_stride, fs, nframes, width, height, C.byref(params),
                                        pattern_data, stream))

    def videomark_box_sums(self, luma, pixel_stride, row_stride, width, height, params, nframes=1, frame_stride=None, stream=None):
        """-> uint64 array [nframes, n_boxes]"""
        fs = frame_stride if frame_stride is not None else row_stride * height
        d = DeviceBuffer(self, nframes * VIDEOMARK_MAX_BOXES * 8)
        n = C.c_int(0)
        check(lib.b200vf_videomark_box_sums(self.h, _ptr(luma), pixel_stride, row_stride, fs, nframes, width, height, C.byref(params),
                                            d.ptr, C.byref(n), stream))
        out = self.download(d, dtype=np.uint64, stream=stream).reshape(nframes, VIDEOMARK_MAX_BOXES)[:, :n.value]
        d.free()
        return out

    def smooth_plane(self, src, dst, stride, width, height, tolerance=8, filtersize=3, nframes=1, frame_stride=None, stream=None):
        fs = frame_stride if frame_stride is not None else stride * height
        check(lib.b200vf_smooth_plane(self.h, _ptr(src), stride, fs, _ptr(dst), stride, fs, width, height, nframes,
                                      tolerance, filtersize, stream))

    def gauss_selftest_div1(self, divisor, e, lo_bits, hi_bits):
        """mismatches between RN(a + a*e) and IEEE a / divisor over fp32 bit patterns (the streaming blur's division)"""
        bad = C.c_ulonglong(0)
        check(lib.b200vf_gauss_selftest_div1(self.h, float(divisor), float(e), int(lo_bits), int(hi_bits), C.byref(bad)))
        return int(bad.value)

    def gauss_selftest_finish(self, lo_bits=0, hi_bits=0xffffffff):
        """mismatches between the blur's fp32-only final rounding and (guint8) CLAMP (q + 0.5 [fp64], 0, 255)"""
        bad = C.c_ulonglong(0)
        check(lib.b200vf_gauss_selftest_finish(self.h, int(lo_bits), int(hi_bits), C.byref(bad)))
        return bad.value

    def coloreffects_rgb(self, data, width, height, row_stride, pixel_stride, offs, table, map_luma, nframes=1,
                         stream=None):
        t = np.ascontiguousarray(table, np.uint8)
        check(lib.b200vf_coloreffects_rgb(self.h, _ptr(data), width, height, row_stride, row_stride * height, nframes,
                                          pixel_stride, offs[0], offs[1], offs[2], _hptr(t), int(map_luma), stream))

    def coloreffects_ayuv(self, data, width, height, row_stride, offs, table, map_luma, nframes=1, stream=None):
        t = np.ascontiguousarray(table, np.uint8)
        check(lib.b200vf_coloreffects_ayuv(self.h, _ptr(data), width, height, row_stride, row_stride * height, nframes,
                                           offs[0], offs[1], offs[2], _hptr(t), int(map_luma), stream))

    def chromahold(self, data, width, height, row_stride, offs, target, tolerance, nframes=1, stream=None):
        check(lib.b200vf_chromahold(self.h, _ptr(data), width, height, row_stride, row_stride * height, nframes,
                                    offs[0], offs[1], offs[2], target[0], target[1], target[2], tolerance, stream))

    def remap(self, src, dst, index, width, height, pixel_stride, row_stride, fill=0, nframes=1, stream=None):
        check(lib.b200vf_remap(self.h, _ptr(src), _ptr(dst), _ptr(index), width, height, pixel_stride, row_stride,
                               row_stride * height, nframes, fill, stream))

    def diffuse(self, src, dst, width, height, pixel_stride, row_stride, sin_table, cos_table, off_edge=1, fill=0, seed=0,
                first_frame=0, nframes=1, first_row=0, full_height=None, stream=None):
        full_height = height if full_height is None else full_height
        st = np.ascontiguousarray(sin_table, np.float64)
        ct = np.ascontiguousarray(cos_table, np.float64)
        check(lib.b200vf_diffuse(self.h, _ptr(src), _ptr(dst), width, height, first_row, full_height, pixel_stride, row_stride,
                                 row_stride * full_height, row_stride * height, nframes, _hptr(st), _hptr(ct), off_edge, fill,
                                 seed, first_frame, stream))

    def remap_packed(self, src, dst, packed, width, height, fill=0, nframes=1, stream=None):
        check(lib.b200vf_remap_packed(self.h, _ptr(src), _ptr(dst), _ptr(packed), width, height, 4 * width * height,
                                      nframes, fill, stream))

    def element(self, factory):
        return Element(self, factory)


MAP_READ, MAP_WRITE, MAP_DEVICE = 1, 2, 4
NEED_UPLOAD, NEED_DOWNLOAD = 1, 2


class Memory:
    """b200vf_memory: device storage + lazy pinned staging + transfer flags (the GstMemory of the HBM pool)."""

    def __init__(self, ctx, nbytes=None, handle=None):
        if handle is None:
            h = _vp()
            check(lib.b200vf_memory_new(ctx.h, nbytes, C.byref(h)))
            handle = h
        self.h, self.ctx = handle, ctx
        ctx._children.add(self)

    @property
    def nbytes(self):
        return int(lib.b200vf_memory_size(self.h))

    @property
    def flags(self):
        return int(lib.b200vf_memory_flags(self.h))

    @property
    def pending_stages(self):
        return int(lib.b200vf_memory_pending_stages(self.h))

    def map(self, flags, stream=None):
        p = _vp()
        check(lib.b200vf_memory_map(self.h, flags, C.byref(p), stream))
        return p.value

    def unmap(self):
        check(lib.b200vf_memory_unmap(self.h))

    def write(self, array, stream=None):
        """host writer: map WRITE, fill the pinned staging buffer, unmap (marks NEED_UPLOAD; no copy happens yet)"""
        a = np.ascontiguousarray(array).view(np.uint8).reshape(-1)
        assert a.size <= self.nbytes
        p = self.map(MAP_WRITE, stream)
        C.memmove(p, _hptr(a), a.size)
        self.unmap()

    def read(self, nbytes=None, stream=None):
        """host reader: map READ (launches a pending chain, downloads if the device copy is newer), copy out, unmap"""
        n = self.nbytes if nbytes is None else nbytes
        p = self.map(MAP_READ, stream)
        out = np.empty(n, np.uint8)
        C.memmove(_hptr(out), p, n)
        self.unmap()
        return out

    def close(self):
        if self.h:
            lib.b200vf_memory_unref(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.close()
        except Exception:
            pass


class Pool:
    """b200vf_pool: one HBM slab of equally sized buffers (the GstBufferPool replacement)."""

    def __init__(self, ctx, buf_bytes, n_bufs):
        h = _vp()
        check(lib.b200vf_pool_create(ctx.h, buf_bytes, n_bufs, C.byref(h)))
        self.h, self.ctx = h, ctx
        self._mems = weakref.WeakSet()
        ctx._children.add(self)

    def acquire_memory(self):
        h = _vp()
        check(lib.b200vf_pool_acquire_memory(self.h, C.byref(h)))
        m = Memory(self.ctx, handle=h)
        self._mems.add(m)
        return m

    def close(self):
        if self.h:
            for m in list(self._mems):
                m.close()
            lib.b200vf_pool_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.close()
        except Exception:
            pass


class DeviceBuffer:
    def __init__(self, ctx, nbytes):
        p = _vp()
        check(lib.b200vf_malloc(ctx.h, nbytes, C.byref(p)))
        self.ctx, self.ptr, self.nbytes = ctx, p.value, nbytes
        ctx._children.add(self)

    def free(self):
        if self.ptr and self.ctx.h:
            lib.b200vf_free(self.ctx.h, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---------------------------------------------------------------- host helpers
def lut_burn(adjustment=175):
    a = np.zeros((4, 256), np.uint8)
    check(lib.b200vf_lut_burn(adjustment, _hptr(a)))
    return a


def lut_dodge():
    a = np.zeros((4, 256), np.uint8)
    check(lib.b200vf_lut_dodge(_hptr(a)))
    return a


def lut_chromium(edge_a=200, edge_b=1):
    a = np.zeros((4, 256), np.uint8)
    check(lib.b200vf_lut_chromium(edge_a, edge_b, _hptr(a)))
    return a


def lut_solarize(threshold=127, start=50, end=185):
    a = np.zeros((4, 256), np.uint8)
    check(lib.b200vf_lut_solarize(threshold, start, end, _hptr(a)))
    return a


def lut_compose(first, second):
    a = np.zeros((4, 256), np.uint8)
    f, s = _lut_arg(first), _lut_arg(second)
    check(lib.b200vf_lut_compose(_hptr(f), _hptr(s), _hptr(a)))
    return a


def gaussblur_halo_rows(windowsize, p0, stride, width):
    n = lib.b200vf_gaussblur_halo_rows(windowsize, p0, stride, width)
    if n < 0:
        check(n)
    return n


def gauss_div1_constant(divisor):
    """e with a / divisor == RN(a + a*e) over the blur's dividend range, or None when the divisor is not whitelisted"""
    e = C.c_float(0)
    rc = lib.b200vf_gauss_div1_constant(C.c_float(divisor), C.byref(e))
    return np.float32(e.value) if rc == 0 else None


def gauss_kernel(sigma):
    k = np.zeros(128, np.float32)
    s = np.zeros(128, np.float32)
    ws = lib.b200vf_gauss_kernel(C.c_float(sigma), _hptr(k), _hptr(s), 128)
    if ws < 0:
        check(ws)
    return k[:ws].copy(), s[:ws].copy()


def coloreffects_table(preset):
    """preset: 1 heat, 2 sepia, 3 xray, 4 xpro, 5 yellowblue -> (uint8[768], map_luma)"""
    p = _vp()
    ml = _i(0)
    check(lib.b200vf_coloreffects_table(preset, C.byref(p), C.byref(ml)))
    arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(768,)).copy()
    return arr, ml.value


def gt_build_map(element, width, height, props=None):
    props = props or {}
    names = [k.replace("_", "-").encode() for k in props]
    n = len(names)
    cn = (C.c_char_p * max(n, 1))(*names)
    cv = (C.c_double * max(n, 1))(*[float(v) for v in props.values()])
    m = np.zeros((height, width, 2), np.float64)
    check(lib.b200vf_gt_build_map(element.encode(), width, height, cn, cv, n, _hptr(m)))
    return m


def diffuse_draw(seed, frame, pixel):
    """(angle 0..255, distance in [0, 1)) pixel `pixel` of frame `frame` draws: the generator of csrc/diffuse.cu on the host"""
    a, d = _i(), C.c_double()
    lib.b200vf_diffuse_draw(seed, frame, pixel, C.byref(a), C.byref(d))
    return a.value, d.value


def diffuse_tables(scale):
    s, c = np.zeros(256, np.float64), np.zeros(256, np.float64)
    check(lib.b200vf_diffuse_tables(scale, _hptr(s), _hptr(c)))
    return s, c


def gt_device_map_supported(element):
    """0: no device map; 1: evaluated exactly on the GPU; 2: certified on the GPU, uncertain entries patched in from the host"""
    return lib.b200vf_gt_device_map_supported(element.encode())


def gt_device_last_uncertain():
    """entries the last gt_build_index_device on this thread took from the host's map function"""
    return lib.b200vf_gt_device_last_uncertain()


def gt_build_index_device(ctx, element, width, height, props=None, off_edge=0, stream=None):
    """the int32 gather table built on the GPU (every map but `diffuse`); returns a DeviceBuffer. E_UNSUPPORTED when a
    libm map cannot be certified (too many coordinates on integers): build on the host then"""
    props = props or {}
    names = (C.c_char_p * max(1, len(props)))(*[k.replace("_", "-").encode() for k in props])
    vals = (C.c_double * max(1, len(props)))(*[float(v) for v in props.values()])
    buf = DeviceBuffer(ctx, width * height * 4)
    check(lib.b200vf_gt_build_index_device(ctx.h, element.encode(), width, height, names, vals, len(props), off_edge, buf.ptr, stream))
    return buf


def gt_resolve_map(map_xy, width, height, off_edge):
    idx = np.zeros((height, width), np.int32)
    m = np.ascontiguousarray(map_xy, np.float64)
    check(lib.b200vf_gt_resolve_map(_hptr(m), width, height, off_edge, _hptr(idx)))
    return idx


def gt_pack_index(index, width, height):
    """-> (packed bytes as np.uint8 array, number of raw (uncoded) 128-pixel groups)"""
    a = np.ascontiguousarray(index, np.int32)
    buf = np.zeros(lib.b200vf_gt_packed_bound(width, height), np.uint8)
    used, raw = _sz(0), _sz(0)
    check(lib.b200vf_gt_pack_index(_hptr(a), width, height, _hptr(buf), buf.size, C.byref(used), C.byref(raw)))
    return buf[:used.value].copy(), raw.value


def gt_unpack_index(packed, width, height):
    p = np.ascontiguousarray(packed, np.uint8)
    idx = np.zeros((height, width), np.int32)
    check(lib.b200vf_gt_unpack_index(_hptr(p), p.size, width, height, _hptr(idx)))
    return idx


class Comm:
    """NCCL communicator of the row-shard path (one process per GPU). `bcast(tensor_or_bytes)` is how the caller
    distributes rank 0's 128-byte unique id (torch.distributed, MPI, a file ...)."""

    def __init__(self, ctx, rank, nranks, bcast):
        buf = (C.c_uint8 * 128)()
        if rank == 0:
            check(lib.b200vf_comm_unique_id(buf))
        idb = (C.c_uint8 * 128)(*bcast(bytes(buf)))
        h = _vp()
        check(lib.b200vf_comm_create(ctx.h, idb, rank, nranks, C.byref(h)))
        self.h, self.rank, self.nranks = h, rank, nranks
        ctx._children.add(self)

    def halo_exchange(self, buf, row_bytes, rows, halo, frame_stride, nframes=1, stream=None):
        check(lib.b200vf_comm_halo_exchange(self.h, _ptr(buf), row_bytes, rows, halo, frame_stride, nframes, stream))

    def halo_begin(self, buf, row_bytes, rows, halo, frame_stride, nframes=1, stream=None):
        check(lib.b200vf_comm_halo_begin(self.h, _ptr(buf), row_bytes, rows, halo, frame_stride, nframes, stream))

    def halo_end(self, stream=None):
        check(lib.b200vf_comm_halo_end(self.h, stream))

    def allgather_rows(self, full, row_bytes, full_rows, frame_stride=0, nframes=1, stream=None):
        check(lib.b200vf_comm_allgather_rows(self.h, _ptr(full), row_bytes, full_rows, frame_stride, nframes, stream))

    def exchange_rows(self, full, row_bytes, full_rows, need_lo, need_hi, frame_stride=0, nframes=1, stream=None):
        lo = np.ascontiguousarray(need_lo, np.int32)
        hi = np.ascontiguousarray(need_hi, np.int32)
        check(lib.b200vf_comm_exchange_rows(self.h, _ptr(full), row_bytes, full_rows, _hptr(lo), _hptr(hi), frame_stride,
                                            nframes, stream))

    def close(self):
        if self.h:
            lib.b200vf_comm_destroy(self.h)
            self.h = None


def gt_index_row_range(index, width):
    a = np.ascontiguousarray(index, np.int32)
    lo, hi = _i(0), _i(0)
    check(lib.b200vf_gt_index_row_range(_hptr(a), a.size, width, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def shard_rows(height, rank, nranks):
    r0, r = _i(0), _i(0)
    check(lib.b200vf_shard_rows(height, rank, nranks, C.byref(r0), C.byref(r)))
    return r0.value, r.value


VIDEOMARK_MAX_BOXES = 128


class VideoMarkParams(C.Structure):
    _fields_ = [("pattern_width", C.c_int), ("pattern_height", C.c_int), ("pattern_count", C.c_int), ("pattern_data_count", C.c_int),
                ("left_offset", C.c_int), ("bottom_offset", C.c_int)]

    def __init__(self, pattern_width=4, pattern_height=16, pattern_count=4, pattern_data_count=5, left_offset=0, bottom_offset=0):
        super().__init__(pattern_width, pattern_height, pattern_count, pattern_data_count, left_offset, bottom_offset)


def videomark_detect_decide(params, width, height, row_stride, pixel_stride, sums, center=0.5, sensitivity=0.3, in_pattern=False):
    """-> (in_pattern, message posted, data)"""
    sm = np.ascontiguousarray(sums, np.uint64)
    ip, msg, data = C.c_int(int(in_pattern)), C.c_int(0), C.c_uint64(0)
    check(lib.b200vf_videomark_detect_decide(C.byref(params), width, height, row_stride, pixel_stride, _hptr(sm), center, sensitivity,
                                             C.byref(ip), C.byref(msg), C.byref(data)))
    return bool(ip.value), bool(msg.value), int(data.value)


class FrameLayout(C.Structure):
    _fields_ = [("n_planes", C.c_int), ("offset", C.c_size_t * 4), ("stride", C.c_int * 4), ("row_bytes", C.c_int * 4),
                ("rows", C.c_int * 4)]


class FactoryInfo(C.Structure):
    _fields_ = [(n, C.c_char_p) for n in ("factory", "plugin", "plugin_description", "plugin_license", "type_name",
                                           "parent_type_name", "klass", "long_name", "description", "author")] + \
               [("in_place", _i), ("n_properties", _i), ("n_formats", _i)]


class PropertyInfo(C.Structure):
    _fields_ = [("name", C.c_char_p), ("type", _i), ("min", C.c_double), ("max", C.c_double), ("default", C.c_double),
                ("controllable", _i), ("n_nicks", _i), ("nicks", C.POINTER(C.c_char_p))]


def factories():
    """{factory: {info..., properties: [...], formats: [...]}} - what gst-inspect would print."""
    out = {}
    for i in range(lib.b200vf_factory_count()):
        fi = FactoryInfo()
        check(lib.b200vf_factory_get(i, C.byref(fi)))
        d = {k: (getattr(fi, k).decode() if isinstance(getattr(fi, k), bytes) else getattr(fi, k)) for k, _ in FactoryInfo._fields_}
        props = []
        for j in range(fi.n_properties):
            pi = PropertyInfo()
            check(lib.b200vf_factory_property(fi.factory, j, C.byref(pi)))
            props.append({"name": pi.name.decode(), "type": pi.type, "min": pi.min, "max": pi.max, "default": pi.default,
                          "controllable": bool(pi.controllable), "nicks": [pi.nicks[k].decode() for k in range(pi.n_nicks)]})
        d["properties"] = props
        d["formats"] = [lib.b200vf_factory_format(fi.factory, j).decode() for j in range(fi.n_formats)]
        out[d["factory"]] = d
    return out


class Element:
    """Mirror of a reference element: factory name, properties, caps, transform."""

    def __init__(self, ctx, factory):
        h = _vp()
        check(lib.b200vf_element_factory_make(ctx.h if ctx is not None else None, factory.encode(), C.byref(h)))
        self.h, self.ctx, self.factory = h, ctx, factory
        if ctx is not None:
            ctx._children.add(self)

    def close(self):
        if self.h:
            lib.b200vf_element_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_property(self, name, value):
        if isinstance(value, str):
            check(lib.b200vf_element_set_property_string(self.h, name.encode(), value.encode()))
        else:
            check(lib.b200vf_element_set_property(self.h, name.encode(), float(value)))

    def get_property(self, name):
        v = C.c_double(0)
        check(lib.b200vf_element_get_property(self.h, name.encode(), C.byref(v)))
        return v.value

    def set_caps(self, in_format, out_format, width, height):
        check(lib.b200vf_element_set_caps(self.h, in_format.encode(), (out_format or in_format).encode(), width, height))

    def unit_size(self):
        a, b = _sz(0), _sz(0)
        check(lib.b200vf_element_unit_size(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def transform(self, frames_in, nframes=1):
        """Host path: numpy in -> numpy out (upload, kernel, download inside)."""
        a = np.ascontiguousarray(frames_in, np.uint8)
        in_b, out_b = self.unit_size()
        assert a.nbytes == in_b * nframes, (a.nbytes, in_b, nframes)
        out = np.empty(out_b * nframes, np.uint8)
        check(lib.b200vf_element_transform_host(self.h, _hptr(a), _hptr(out), nframes))
        return out

    def default_layout(self, side):
        lay = FrameLayout()
        check(lib.b200vf_element_default_layout(self.h, side, C.byref(lay)))
        return lay

    def transform_layout(self, buf_in, lay_in, buf_out, lay_out):
        """one frame whose planes lie at lay.offset[i] / lay.stride[i] inside the numpy buffers (GstVideoMeta layouts)"""
        check(lib.b200vf_element_transform_host_layout(self.h, buf_in.ctypes.data, C.byref(lay_in) if lay_in is not None else None,
                                                       buf_out.ctypes.data, C.byref(lay_out) if lay_out is not None else None))

    def set_host_mode(self, mode):
        check(lib.b200vf_element_set_host_mode(self.h, mode))

    def transform_host_ptr(self, h_in, h_out, nframes):
        check(lib.b200vf_element_transform_host(self.h, h_in, h_out, nframes))

    def transform_mem(self, m_in, m_out, nframes=1, stream=None):
        """the transform vfunc on Memory objects: frames stay in HBM, per-pixel elements join pending chains"""
        check(lib.b200vf_element_transform(self.h, m_in.h, m_out.h, nframes, stream))

    def transform_device(self, d_in, d_out, nframes=1, stream=None):
        check(lib.b200vf_element_transform_device(self.h, _ptr(d_in), _ptr(d_out), nframes, stream))

    def set_rng_seed(self, seed, next_frame=0):
        check(lib.b200vf_element_set_rng_seed(self.h, seed, next_frame))

    def rng_state(self):
        a, b = C.c_uint64(), C.c_uint64()
        check(lib.b200vf_element_get_rng_state(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_values(self):
        v = (C.c_double * 8)()
        n = lib.b200vf_element_last_values(self.h, v, 8)
        if n < 0:
            check(n)
        return [v[i] for i in range(n)]

    def last_events(self):
        """scenechange: per frame of the last transform call, True where the reference pushes its force-key-unit event"""
        n = lib.b200vf_element_last_events(self.h, None, 0)
        if n < 0:
            check(n)
        flags = (C.c_int * max(n, 1))()
        lib.b200vf_element_last_events(self.h, flags, n)
        return [bool(flags[i]) for i in range(n)]


def videoanalyse_finish(total, total_sq, width, height):
    """(luma-average, luma-variance) of the videoanalyse element from the two moments (host arithmetic)"""
    a, v = C.c_double(0), C.c_double(0)
    check(lib.b200vf_videoanalyse_finish(int(total), int(total_sq), width, height, C.byref(a), C.byref(v)))
    return a.value, v.value


class SceneChange:
    """the scenechange element's decision state (gstscenechange.c:196-236) on the library's host code"""

    class _State(C.Structure):
        _fields_ = [("diffs", C.c_double * 5), ("n_diffs", C.c_int)]

    def __init__(self):
        self.st = self._State()
        check(lib.b200vf_scenechange_reset(C.byref(self.st)))

    def update(self, score):
        ch = C.c_int(0)
        check(lib.b200vf_scenechange_update(C.byref(self.st), float(score), C.byref(ch)))
        return bool(ch.value)
