/* b200vf.h - C-ABI of the B200-native per-pixel video-filter hot path.
 *
 * This is the drop-in boundary: a GLib-free, torch-free
 * `extern "C"` surface of plain pointers and sizes that the element shells
 * of gst-plugins-bad's `bayer`, `gaudieffects`, `coloreffects` and
 * `geometrictransform` plugins call from their transform vfuncs instead of
 * their ORC / scalar C inner loops.  Each entry point names the reference
 * interface it replaces (paths relative to the gst-plugins-bad 1.19.2 tree).
 *
 * Conventions
 *   - every function returns B200VF_OK (0) or a negative b200vf_status; nothing
 *     aborts; b200vf_last_error() gives a thread-local message;
 *   - `d_` pointers are device (HBM) addresses on the context's GPU;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the context's own
 *     stream); all work is asynchronous and stream-ordered, no call
 *     synchronises unless it says so;
 *   - `nframes`/`*_frame_stride`: every op takes a batch of equally shaped
 *     frames laid out at a fixed byte pitch (how the HBM pool lays them out),
 *     one launch for the whole batch; nframes = 1 for a single GstBuffer;
 *   - there is NO CPU fallback: without an sm_100 device b200vf_ctx_create
 *     fails with B200VF_E_NO_DEVICE and nothing else can be called.
 */
#ifndef B200VF_H
#define B200VF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VF_VERSION_MAJOR 0
#define B200VF_VERSION_MINOR 1

typedef enum b200vf_status {
  B200VF_OK = 0,
  B200VF_E_INVAL = -1,        /* bad argument / outside the reference's domain */
  B200VF_E_NO_DEVICE = -2,    /* no sm_100 GPU (the product has no CPU path)   */
  B200VF_E_CUDA = -3,         /* a CUDA runtime/driver call failed             */
  B200VF_E_NOMEM = -4,
  B200VF_E_UNSUPPORTED = -5,  /* caps/format the reference element rejects too */
  B200VF_E_NOT_NEGOTIATED = -6, /* GST_FLOW_NOT_NEGOTIATED analogue            */
  B200VF_E_NCCL = -7,
  B200VF_E_PROPERTY = -8      /* unknown property / out-of-range value         */
} b200vf_status;

typedef struct b200vf_ctx b200vf_ctx;
typedef struct b200vf_pool b200vf_pool;
typedef struct b200vf_element b200vf_element;
typedef struct b200vf_comm b200vf_comm;
typedef struct b200vf_memory b200vf_memory;

/* ------------------------------------------------------------------ context */
int b200vf_version (void);                       /* major*100 + minor */
const char *b200vf_last_error (void);
const char *b200vf_status_string (int status);

/* Binds to CUDA device `device` (must be compute capability 10.x), creates the
 * context's stream. Mirrors the "register nothing if the driver is missing"
 * rule of sys/nvcodec/plugin.c:72-103. */
int b200vf_ctx_create (int device, b200vf_ctx **out);
void b200vf_ctx_destroy (b200vf_ctx *ctx);
int b200vf_ctx_device (const b200vf_ctx *ctx);
void *b200vf_ctx_stream (const b200vf_ctx *ctx);
int b200vf_ctx_sm_count (const b200vf_ctx *ctx);
int b200vf_ctx_synchronize (b200vf_ctx *ctx, void *stream);
/* Number of kernels this library has launched on this context so far (the
 * bench's `gpu_launches` claim is the difference across the timed region). */
uint64_t b200vf_ctx_launch_count (const b200vf_ctx *ctx);
/* Name of the kernel variant the last call on this context dispatched to
 * (e.g. "bayer2rgb_tma", "bayer2rgb_direct"); for tests and profiles. */
const char *b200vf_ctx_last_kernel (const b200vf_ctx *ctx);
/* Force a kernel variant for A/B measurements: 0 auto, 1 direct (no TMA), 2 TMA. */
int b200vf_ctx_set_variant (b200vf_ctx *ctx, int variant);

/* ---------------------------------------------------------- HBM buffer pool
 * Replaces the default system-memory GstVideoBufferPool the elements get today
 * (none of them overrides propose/decide_allocation); modelled on
 * sys/nvcodec/gstcudabufferpool.c:55-222 + gstcudamemory.c:95-407: device
 * storage is primary, a pinned host staging buffer is created lazily.
 * Buffers have `buf_bytes` usable bytes + >= 64 zeroed slack bytes (D5). */
int b200vf_pool_create (b200vf_ctx *ctx, size_t buf_bytes, int n_bufs, b200vf_pool **out);
void b200vf_pool_destroy (b200vf_pool *pool);
int b200vf_pool_acquire (b200vf_pool *pool, int *buf_index);      /* B200VF_E_NOMEM when drained */
int b200vf_pool_release (b200vf_pool *pool, int buf_index);
void *b200vf_pool_device_ptr (b200vf_pool *pool, int buf_index);
void *b200vf_pool_host_ptr (b200vf_pool *pool, int buf_index);   /* pinned staging, lazily allocated */
size_t b200vf_pool_buf_bytes (const b200vf_pool *pool);
size_t b200vf_pool_buf_pitch (const b200vf_pool *pool);           /* bytes between consecutive buffers */
int b200vf_pool_upload (b200vf_pool *pool, int buf_index, const void *host_src, size_t bytes, void *stream);
int b200vf_pool_download (b200vf_pool *pool, int buf_index, void *host_dst, size_t bytes, void *stream);
/* Raw helpers for callers that own their memory. */
int b200vf_malloc (b200vf_ctx *ctx, size_t bytes, void **d_out);  /* zero-filled, +64 B slack */
int b200vf_free (b200vf_ctx *ctx, void *d_ptr);
int b200vf_host_alloc (size_t bytes, void **h_out);               /* pinned */
int b200vf_host_free (void *h_ptr);
int b200vf_memcpy_h2d (b200vf_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, void *stream);
int b200vf_memcpy_d2h (b200vf_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, void *stream);

/* ------------------------------------------------------- device memory object
 * What a GstMemory of the HBM pool is (modelled on GstCudaMemory, sys/nvcodec/gstcudamemory.c:
 * 95-154 alloc, 257-407 map/unmap/transfer flags): device storage is primary, a pinned host staging buffer appears on
 * the first host map, and two flags say which copy is stale. Elements take memories (b200vf_element_transform), so a
 * chain of elements keeps its frames in HBM: the only transfers are the upload of what a host writer staged and the
 * download when a host reader maps.
 *   map (B200VF_MAP_DEVICE | ...): the HBM address, after uploading staged bytes (NEED_UPLOAD); WRITE marks the
 *     staging copy stale (NEED_DOWNLOAD)                         - cuda_mem_map with GST_MAP_CUDA, :331-356
 *   map (READ / WRITE without DEVICE): the pinned staging address, after downloading (NEED_DOWNLOAD) and waiting for
 *     the stream; unmap of a WRITE map marks the device copy stale (NEED_UPLOAD)     - :357-407
 * Deferred per-pixel chains: a memory may hold a PENDING chain instead of bytes - bayer2rgb, then at most one
 * luma-mapped coloreffects preset, then any number of per-channel LUT elements (burn, dodge, chromium, solarize,
 * coloreffects' per-channel presets); or LUT elements alone. The chain is launched as ONE kernel
 * (b200vf_bayer2rgb_fused / b200vf_lut4 with the LUTs composed on the host) when the bytes are needed: a map, or an
 * element that cannot join the chain. `bayer2rgb ! coloreffects ! solarize` (BASELINE.json configs[4]) through
 * memories is 1 upload, 1 launch, 1 download. The chain holds a reference to its source memory.
 * Memories are reference counted (GstMiniObject): a pool memory returns to its pool when the last reference goes. */
/* Streams: every map / element call names the stream it works on (NULL = the context's). A memory remembers the
 * stream of its last device use; a different stream is ordered behind it with an event before it touches the bytes. */
#define B200VF_MAP_READ 1
#define B200VF_MAP_WRITE 2
#define B200VF_MAP_DEVICE 4
#define B200VF_MEMORY_NEED_UPLOAD 1
#define B200VF_MEMORY_NEED_DOWNLOAD 2
int b200vf_memory_new (b200vf_ctx *ctx, size_t bytes, b200vf_memory **out);     /* zero-filled, + 64 B slack (D5) */
int b200vf_pool_acquire_memory (b200vf_pool *pool, b200vf_memory **out);        /* B200VF_E_NOMEM when drained */
b200vf_memory *b200vf_memory_ref (b200vf_memory *mem);
void b200vf_memory_unref (b200vf_memory *mem);
size_t b200vf_memory_size (const b200vf_memory *mem);
unsigned b200vf_memory_flags (const b200vf_memory *mem);
int b200vf_memory_is_writable (const b200vf_memory *mem);                       /* refcount == 1 (gst_mini_object_is_writable) */
int b200vf_memory_pending_stages (const b200vf_memory *mem);                    /* elements recorded, not yet launched */
int b200vf_memory_map (b200vf_memory *mem, int flags, void **data, void *stream);
int b200vf_memory_unmap (b200vf_memory *mem);
/* host <-> device copies made through memories on this context so far (tests, bench `e2e`) */
int b200vf_ctx_transfer_counts (const b200vf_ctx *ctx, uint64_t *h2d_count, uint64_t *h2d_bytes, uint64_t *d2h_count,
    uint64_t *d2h_bytes);

/* -------------------------------------------------------------- bayer plugin
 * b200vf_bayer2rgb replaces gst_bayer2rgb_process (gst/bayer/gstbayer2rgb.c:
 * 387-451) and the ORC programs it drives (gstbayerorc.orc:3-248).
 *   pattern: 0 bggr, 1 gbrg, 2 grbg, 3 rggb   (enum gstbayer2rgb.c:95-101)
 *   r_off/g_off/b_off: GST_VIDEO_INFO_COMP_OFFSET of the negotiated src format
 *     (gstbayer2rgb.c:269-271); must be one of the four triples the reference
 *     dispatches on (:409-421): (2,1,0) (3,2,1) (1,2,3) (0,1,2); the remaining
 *     byte is written 255.
 *   src_stride is GST_ROUND_UP_4(width) in the element (:477); dst_stride 4*width.
 *   Domain: even width >= 4, height >= 3 (else B200VF_E_INVAL; the reference
 *   reads uninitialised memory there). Edge rules reproduced bit-exactly: top
 *   mirrors row 1, bottom uses row height-4, right edge copies (:372-380). */
int b200vf_bayer2rgb (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int height, int nframes,
    int pattern, int r_off, int g_off, int b_off, void *stream);

/* Row-sharded variant (multi-GPU): this rank owns global rows
 * [row0, row0+rows) of a frame of `full_height` rows; d_src points at the
 * shard's first row and must be preceded by one valid halo row and followed by
 * one (the rows b200vf_comm_halo_exchange fills); the global edge rules are
 * applied with global row indices, the bottom rule (row height-4) reads
 * d_bottom_m4 = pointer to global row full_height-4 if this rank owns the
 * last row (it lies inside the last shard when that shard has >= 4 rows). */
int b200vf_bayer2rgb_shard (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int full_height, int row0, int rows,
    int nframes, int pattern, int r_off, int g_off, int b_off, void *stream);

/* b200vf_bayer2rgb_shard with the fused epilogue of b200vf_bayer2rgb_fused (BASELINE.json config 5 on N GPUs). */
int b200vf_bayer2rgb_shard_fused (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int full_height, int row0, int rows,
    int nframes, int pattern, int r_off, int g_off, int b_off, const uint8_t *luma_table768,
    const uint8_t lut[4][256], void *stream);

/* Replaces the per-pixel select loop of gst_rgb2bayer_transform
 * (gst/bayer/gstrgb2bayer.c:254-267); src is ARGB (4 B/px). */
int b200vf_rgb2bayer (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int height, int nframes,
    int pattern, void *stream);

/* ------------------------------------------------------- gaudieffects plugin
 * Per-byte-position LUT over packed 4-byte pixels: out.byte[c] = lut[c][in.byte[c]].
 * Replaces gaudi_orc_burn (gstgaudieffectsorc.orc:1-25) and the static
 * transform() loops of gstdodge.c:231-254, gstchromium.c:282-338,
 * gstsolarize.c:286-339 (all pure per-channel functions) and the per-channel
 * presets of coloreffects. LUT builders below reproduce each element's
 * arithmetic on the host (libm cos for chromium stays on the host, §8c-ii).
 * npix = width*height (the reference loops flat, stride == 4*width). */
int b200vf_lut4 (b200vf_ctx *ctx, const uint8_t *d_src, uint8_t *d_dst, size_t npix_total,
    const uint8_t lut[4][256], void *stream);
int b200vf_lut_burn (int adjustment, uint8_t lut[4][256]);                 /* gstburn.c:214-250 */
int b200vf_lut_dodge (uint8_t lut[4][256]);                                /* gstdodge.c:231-254 */
int b200vf_lut_chromium (int edge_a, int edge_b, uint8_t lut[4][256]);     /* gstchromium.c:282-338 */
int b200vf_lut_solarize (int threshold, int start, int end, uint8_t lut[4][256]); /* gstsolarize.c:286-339 */
/* lut_out[c][v] = second[c][first[c][v]] : host-side fusion of consecutive per-channel elements */
int b200vf_lut_compose (const uint8_t first[4][256], const uint8_t second[4][256], uint8_t lut_out[4][256]);

/* gstexclusion.c:256-284 (red uses green*red, :269-270); factor in [1,175]. */
int b200vf_exclusion (b200vf_ctx *ctx, const uint8_t *d_src, uint8_t *d_dst, size_t npix_total,
    int factor, void *stream);

/* gstdilate.c:258-345; frames of width x height u32 pixels, stride 4*width.
 * d_below (may be NULL) = the row under the last row when the frame is a row
 * shard (else the last row's `down` neighbour is itself). */
int b200vf_dilate (b200vf_ctx *ctx, const uint8_t *d_src, uint8_t *d_dst, int width, int height,
    size_t frame_stride, int nframes, int erode, const uint8_t *d_below, void *stream);

/* gaussianblur. b200vf_gauss_kernel = make_gaussian_kernel (gstgaussblur.c:
 * 361-422) on the host (libm pow/sqrt); returns windowsize (odd, <= 101) or <0.
 * b200vf_gaussblur = gaussian_smooth + blur_row_x (:259-356) with the element's
 * gst_video_frame_copy (:252) folded in: d_dst receives the whole frame, bytes
 * [p0, p0 + 4*width*height) blurred (p0 = COMP_OFFSET of component 0: 1 for
 * AYUV), other bytes copied. exact != 0: separate fp32 multiply and
 * add in the reference's tap order (bit-exact); exact == 0: fused multiply-add
 * (faster, within 1 LSB of the u8 output). Row-shard arguments: the frame is
 * rows [row0,row0+rows) of full_height; halo rows (center above/below; center+1
 * when p0 > 0 and stride == 4*width, because a pixel's last p0 bytes are the
 * first bytes of the next row) must be present around d_src unless at the
 * global edge. */
int b200vf_gauss_kernel (float sigma, float *kernel, float *kernel_sum, int capacity);
/* halo rows (above and below) a row shard must be given: windowsize/2, +1 when
 * p0 > 0 and stride == 4*width (see above); <0 = status */
int b200vf_gaussblur_halo_rows (int windowsize, int p0, int stride, int width);
int b200vf_gaussblur (b200vf_ctx *ctx, const uint8_t *d_src, uint8_t *d_dst, int width, int full_height,
    int row0, int rows, int stride, size_t frame_stride, int nframes, int p0,
    const float *kernel, const float *kernel_sum, int windowsize, int exact, void *stream);
/* Test hook: the blur divides by per-column/row constants through their
 * reciprocal plus two FMA corrections; this counts the fp32 values a (bit
 * patterns [lo_bits, hi_bits)) for which that differs from IEEE a / divisor
 * (must be 0 over the range the host enables it for, see gaussblur.cu div_rn). */
int b200vf_gauss_selftest_div (b200vf_ctx *ctx, float divisor, uint32_t lo_bits, uint32_t hi_bits,
    unsigned long long *mismatches);
/* Test hooks of the streaming blur's interior division: the full kernel sum b that make_gaussian_kernel produces is
 * within a few ulp of 1.0, and for a whitelist of such b the quotient a / b equals RN (a + a * e) - one FMA - for
 * a == 0 and every fp32 a in [2^-64, 2^13]. b200vf_gauss_div1_constant returns the whitelisted e for b
 * (B200VF_E_UNSUPPORTED: not whitelisted, the blur then takes the general kernel); b200vf_gauss_selftest_div1
 * counts the bit patterns a in [lo_bits, hi_bits) for which RN (a + a * e) differs from IEEE a / divisor. */
int b200vf_gauss_div1_constant (float divisor, float *e_out);
int b200vf_gauss_selftest_div1 (b200vf_ctx *ctx, float divisor, float e, uint32_t lo_bits, uint32_t hi_bits,
    unsigned long long *mismatches);
/* Test hook: the blur's last step, (guint8) CLAMP (q + 0.5 [double], 0, 255)
 * (gstgaussblur.c:348-351), is computed without fp64; this counts the fp32 bit
 * patterns q in [lo_bits, hi_bits] (inclusive) for which it differs from the
 * fp64 expression (0 .. 0xffffffff covers all of fp32; must be 0). */
int b200vf_gauss_selftest_finish (b200vf_ctx *ctx, uint32_t lo_bits, uint32_t hi_bits,
    unsigned long long *mismatches);

/* ------------------------------------------------------- coloreffects plugin
 * In place (transform_frame_ip, gstcoloreffects.c:479-501).
 * table: one of the five 256x3 preset tables (b200vf_coloreffects_table);
 * map_luma as set per preset (:503-548); offs = COMP_POFFSET of R,G,B (or
 * Y,U,V); pixel_stride 3 or 4; row_stride in bytes.
 * b200vf_coloreffects_rgb replaces gst_color_effects_transform_rgb (:303-359),
 * b200vf_coloreffects_ayuv replaces gst_color_effects_transform_ayuv (:361-435). */
int b200vf_coloreffects_table (int preset, const uint8_t **table768, int *map_luma); /* 1 heat..5 yellowblue */
int b200vf_coloreffects_rgb (b200vf_ctx *ctx, uint8_t *d_data, int width, int height, int row_stride,
    size_t frame_stride, int nframes, int pixel_stride, int off_r, int off_g, int off_b,
    const uint8_t *table768, int map_luma, void *stream);
int b200vf_coloreffects_ayuv (b200vf_ctx *ctx, uint8_t *d_data, int width, int height, int row_stride,
    size_t frame_stride, int nframes, int off_y, int off_u, int off_v,
    const uint8_t *table768, int map_luma, void *stream);
/* gst_chroma_hold_process_xrgb (gstchromahold.c:317-360), in place, 4 B/px. */
int b200vf_chromahold (b200vf_ctx *ctx, uint8_t *d_data, int width, int height, int row_stride,
    size_t frame_stride, int nframes, int off_r, int off_g, int off_b,
    int target_r, int target_g, int target_b, int tolerance, void *stream);

/* -------------------------------------------------- geometrictransform plugin
 * The reference precomputes a gdouble (x,y) map per output pixel with libm on
 * the CPU (gst_geometric_transform_generate_map, gstgeometrictransform.c:80-128)
 * and resolves it per frame in do_map (:167-207). Here the host resolves the
 * map ONCE into a compact int32 source-pixel index per output pixel
 * (ty*width+tx, or -1 = keep the fill value) with exactly do_map's off-edge
 * policy and truncation, and the kernel gathers.
 *   element: "fisheye" "circle" ... (the 16 factory names of plugin.c:40-62,
 *   `diffuse` excluded: it has no precalculated map, see b200vf_diffuse). props: element properties as name/value
 *   pairs ("x-center", "zoom", ...; enum/int properties as doubles).
 *   off_edge: 0 ignore, 1 clamp, 2 wrap (enum :57-75). */
int b200vf_gt_build_map (const char *element, int width, int height, const char *const *prop_names,
    const double *prop_values, int nprops, double *map_xy /* [height][width][2] */);
int b200vf_gt_resolve_map (const double *map_xy, int width, int height, int off_edge, int32_t *index_out);
/* The same table built ON the GPU: d_index equals what b200vf_gt_build_map + b200vf_gt_resolve_map give - without the
 * host rebuild (0.2 s at 8K) and the 132 MB upload every time a GstController moves a property (needs_remap,
 * gstgeometrictransform.c:256-263). b200vf_gt_device_map_supported says how:
 *   1  maps whose arithmetic is +, -, *, /, sqrt, comparisons and table lookups only (mirror, square, stretch, bulge,
 *      tunnel, perspective, marble): the kernel evaluates the reference's fp64 expressions operation for operation
 *      (compiled without FMA contraction); asynchronous on `stream`;
 *   2  maps that call libm (fisheye, circle, kaleidoscope, pinch, rotate, sphere, twirl, waterripple): the kernel
 *      evaluates with CUDA's libm and certifies every entry whose coordinates are further from an integer than a bound
 *      on the disagreement with glibc; the others are evaluated by the host's map function and patched in. The call
 *      synchronises `stream`. When more than 1/64 of the entries are uncertain (rotate at angle 0) it returns
 *      B200VF_E_UNSUPPORTED and the caller builds the table on the host;
 *   0  no table for this element (`diffuse` draws per frame: b200vf_diffuse).
 * b200vf_gt_device_last_uncertain: how many entries the last build on this thread took from the host. */
int b200vf_gt_device_map_supported (const char *element);
long long b200vf_gt_device_last_uncertain (void);
int b200vf_gt_build_index_device (b200vf_ctx *ctx, const char *element, int width, int height,
    const char *const *prop_names, const double *prop_values, int nprops, int off_edge, int32_t *d_index, void *stream);
/* diffuse (gstdiffuse.c:151-231): the one element without a precalculated map - every pixel of every frame draws an
 * angle (0..255) and a distance in [0, 1) and copies the pixel at (x + distance * sin_table[angle],
 * y + distance * cos_table[angle]) under do_map's policy. The reference draws from GLib's global generator, so no two
 * runs of it agree (parity unpinned by construction); here the draw is a stateless function of (seed, frame number,
 * pixel number = y * width + x), the same on host (b200vf_diffuse_draw) and device, and everything around it - fp64
 * arithmetic with separately rounded multiply and add, policy, truncation, bounds, the cleared frame - is the
 * reference's. b200vf_diffuse_tables = diffuse_prepare (:151-165) with the host's libm. Rows [first_row,
 * first_row + height) of a full_height frame are written (d_dst points at row first_row); d_src is the whole frame.
 * Frame f of the call draws with frame number first_frame + f. */
void b200vf_diffuse_draw (uint64_t seed, uint64_t frame, uint64_t pixel, int *angle, double *distance);
int b200vf_diffuse_tables (double scale, double *sin_table /* [256] */, double *cos_table /* [256] */);
int b200vf_diffuse (b200vf_ctx *ctx, const uint8_t *d_src, uint8_t *d_dst, int width, int height, int first_row,
    int full_height, int pixel_stride, int row_stride, size_t src_frame_stride, size_t dst_frame_stride, int nframes,
    const double *sin_table, const double *cos_table, int off_edge, uint32_t fill, uint64_t seed, uint64_t first_frame,
    void *stream);
/* fill: 32-bit pattern the cleared frame holds (0, or 0x808010ff for AYUV =
 * GST_WRITE_UINT32_BE(0xff108080), :244-252); pixel_stride 1,2,3 or 4. */
int b200vf_remap (b200vf_ctx *ctx, const uint8_t *d_src, uint8_t *d_dst, const int32_t *d_index,
    int width, int height, int pixel_stride, int row_stride, size_t frame_stride, int nframes,
    uint32_t fill, void *stream);
/* Packed index table for 4-byte pixels and contiguous rows: the resolved table
 * re-coded as per-pixel steps of the source position (1.5 B/px instead of 4;
 * chunks of 8 pixels the steps cannot express - ignored pixels, map
 * discontinuities - stay raw). Lossless: b200vf_gt_unpack_index returns the
 * table b200vf_gt_pack_index was given, and b200vf_remap_packed writes what
 * b200vf_remap writes. pack: host -> host buffer of b200vf_gt_packed_bound
 * bytes (*used = bytes to upload, *raw_chunks = 8-pixel chunks left uncoded);
 * width, height <= 32767. Measured (profiles/r02_remap.md): in a BATCH of
 * frames the int32 table is shared through L2 and the packed table runs at
 * 0.93x its frame rate, so the element mirror keeps the int32 table; for
 * single-frame launches the int32 table is read from HBM every frame (12 B/px
 * moved for 8 B/px credited) and the packed form (9.5 B/px) is the better one,
 * as it is for callers that hold many maps resident (2.7x smaller). */
size_t b200vf_gt_packed_bound (int width, int height);
int b200vf_gt_pack_index (const int32_t *index, int width, int height, void *packed, size_t capacity,
    size_t *used, size_t *raw_chunks);
int b200vf_gt_unpack_index (const void *packed, size_t size, int width, int height, int32_t *index);
int b200vf_remap_packed (b200vf_ctx *ctx, const uint8_t *d_src, uint8_t *d_dst, const void *d_packed,
    int width, int height, size_t frame_stride, int nframes, uint32_t fill, void *stream);

/* ---------------------------------------------------- videofiltersbad plugin
 * (sibling per-pixel filters on planar / packed YUV.)
 *
 * zebrastripe: gst_zebra_stripe_transform_frame_ip (gst/videofilters/
 * gstzebrastripe.c:205-253), in place on the luma samples. d_luma = the first
 * luma byte of frame 0, i.e. plane 0 + the format's byte offset (+1 for UYVY
 * and AYUV, :232-237); pixel_stride = GST_VIDEO_FORMAT_INFO_PSTRIDE (finfo, 0):
 * 1 (I420 Y444 Y42B Y41B NV12 NV21 YV12), 2 (YUY2 UYVY), 4 (AYUV). y_threshold
 * = b200vf_zebrastripe_y_threshold (`threshold` property, int [0,100] = 90,
 * :150-151). t = the element's frame counter (:217); frame f of a batch uses
 * t + f. */
int b200vf_zebrastripe_y_threshold (int threshold);
int b200vf_zebrastripe (b200vf_ctx *ctx, uint8_t *d_luma, int pixel_stride, int row_stride, size_t frame_stride,
    int nframes, int width, int height, int y_threshold, int t, void *stream);
/* videodiff: the luma loop of gst_video_diff_transform_frame_ip_planarY
 * (gstvideodiff.c:94-117): out = |new - old| > threshold ? ((i+j+t)&4 ? 16 :
 * 240) : new. threshold = 10 and t = 0 in the element (:89, never counted up).
 * The chroma planes are plain copies (:118-127): cudaMemcpy2DAsync in the
 * shell. Planes and pitches 4-byte aligned (GstVideoInfo pitches are). */
int b200vf_videodiff_luma (b200vf_ctx *ctx, const uint8_t *d_old, int old_stride, size_t old_frame_stride,
    const uint8_t *d_new, int new_stride, size_t new_frame_stride, uint8_t *d_out, int out_stride, size_t out_frame_stride,
    int width, int height, int nframes, int threshold, int t, void *stream);
/* scenechange: orc_sad_nxm_u8 (gstscenechangeorc.orc; C backup
 * gstscenechangeorc-dist.c:147-180) for each of nframes pairs of luma planes:
 * d_sums[f] = sum |a - b| modulo 2^32 (the reference accumulates in
 * orc_uint32). Device result; copy it out after the stream has run.
 * get_frame_score (gstscenechange.c:141-155) = sum / (width * height). */
int b200vf_sad_u8 (b200vf_ctx *ctx, const uint8_t *d_a, int a_stride, size_t a_frame_stride, const uint8_t *d_b,
    int b_stride, size_t b_frame_stride, int width, int height, int nframes, uint32_t *d_sums, void *stream);
/* The decision that follows the score (gstscenechange.c:196-236): host state
 * of SC_N_DIFFS = 5 scores; *change_out = 1 when the element would push its
 * force-key-unit event. reset = the state of a fresh element / after a change. */
#define B200VF_SC_N_DIFFS 5
typedef struct b200vf_scenechange_state { double diffs[B200VF_SC_N_DIFFS]; int n_diffs; } b200vf_scenechange_state;
int b200vf_scenechange_reset (b200vf_scenechange_state *st);
int b200vf_scenechange_update (b200vf_scenechange_state *st, double score, int *change_out);

/* ------------------------------------------------------- videosignal plugin
 * videoanalyse: gst_video_analyse_planar (gst/videosignal/gstvideoanalyse.c:
 * 206-236). b200vf_luma_moments: d_sums[2f] = sum of the luma samples of frame
 * f, d_sums[2f+1] = sum of their squares (uint64, device memory; one pass).
 * b200vf_videoanalyse_finish (host): the element's `luma-average` and
 * `luma-variance` from them, in the reference's arithmetic (integer avg). */
int b200vf_luma_moments (b200vf_ctx *ctx, const uint8_t *d_luma, int stride, size_t frame_stride, int width, int height,
    int nframes, uint64_t *d_sums, void *stream);
int b200vf_videoanalyse_finish (uint64_t sum, uint64_t sum_sq, int width, int height, double *luma_average,
    double *luma_variance);

/* simplevideomark / simplevideomarkdetect (gst/videosignal/gstsimplevideomark.c:348-462, gstsimplevideomarkdetect.c:
 * 420-565). d_luma = the first luma sample (COMP_DATA of component 0), pixel_stride / row_stride = COMP_PSTRIDE /
 * COMP_STRIDE. The boxes - pattern-count calibration boxes alternating black / white, then pattern-data-count boxes
 * spelling pattern-data (most significant bit first) - are walked on the host exactly as the reference walks them
 * (clipping, early exits); at most B200VF_VIDEOMARK_MAX_BOXES are supported.
 *   b200vf_videomark_draw: the mark, in place, one launch per batch.
 *   b200vf_videomark_box_sums: d_sums[f * B200VF_VIDEOMARK_MAX_BOXES + k] = sum of the samples of box k of frame f
 *     (the detector averages the FULL pattern width even where a box is clipped, as the reference does); *n_boxes = how
 *     many boxes the walk visits. Device result: copy it out after the stream has run.
 *   b200vf_videomark_detect_decide (host): the reference's decisions on one frame's sums; *in_pattern is the element's
 *     state, *message = 1 when the element would post its message (have-pattern = *in_pattern afterwards, data = *data). */
#define B200VF_VIDEOMARK_MAX_BOXES 128
typedef struct b200vf_videomark_params {
  int pattern_width, pattern_height, pattern_count, pattern_data_count, left_offset, bottom_offset;
} b200vf_videomark_params;
int b200vf_videomark_draw (b200vf_ctx *ctx, uint8_t *d_luma, int pixel_stride, int row_stride, size_t frame_stride,
    int nframes, int width, int height, const b200vf_videomark_params *params, uint64_t pattern_data, void *stream);
int b200vf_videomark_box_sums (b200vf_ctx *ctx, const uint8_t *d_luma, int pixel_stride, int row_stride, size_t frame_stride,
    int nframes, int width, int height, const b200vf_videomark_params *params, uint64_t *d_sums, int *n_boxes, void *stream);
int b200vf_videomark_detect_decide (const b200vf_videomark_params *params, int width, int height, int row_stride,
    int pixel_stride, const uint64_t *sums, double pattern_center, double pattern_sensitivity, int *in_pattern, int *message,
    uint64_t *data);

/* ------------------------------------------------------------ smooth plugin
 * smooth_filter (gst/smooth/gstsmooth.c:131-176) on one 8-bit plane: the mean
 * (integer division) of the reference sample and of the window samples within
 * +-tolerance of it. Window: columns [x-fs, x+fs], rows [r-fs, r+fs+2] around
 * output row r (the reference reads its reference sample one row above the
 * window centre), clipped as the reference's loop clips them. Rows 0 ..
 * height-2 are written; the last row is never written by the reference and is
 * left untouched here. filtersize <= 8 (negative: empty window, output =
 * input); any tolerance (the reference's int product overflows - undefined -
 * beyond |tolerance| ~ 46000; here every |tolerance| > 255 admits all samples). The element (`active`, `tolerance` 8, `filter-size` 3,
 * `luma-only` TRUE; I420) filters plane 0 and copies or filters planes 1, 2
 * (:178-222). */
int b200vf_smooth_plane (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int height, int nframes, int tolerance,
    int filtersize, void *stream);

/* ------------------------------------------------------------ fused chains
 * bayer2rgb followed by per-channel LUT elements (coloreffects per-channel
 * presets, burn, dodge, chromium, solarize, composed on the host) and/or one
 * luma-mapped coloreffects preset, in ONE kernel: 5 B/px instead of 5+8+8
 * (BASELINE.json config 5). lut (may be NULL) is applied per output byte
 * position; luma_table768 (may be NULL) is a map_luma preset applied first. */
int b200vf_bayer2rgb_fused (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int height, int nframes,
    int pattern, int r_off, int g_off, int b_off, const uint8_t *luma_table768,
    const uint8_t lut[4][256], void *stream);

/* ------------------------------------------------------ multi-GPU row shards
 * One process per GPU. The communicator wraps NCCL (dlopen'd libnccl.so.2, so
 * single-GPU users need no NCCL): the caller distributes the 128-byte unique
 * id from rank 0 by whatever means it has. */
int b200vf_comm_unique_id (uint8_t id_out[128]);
int b200vf_comm_create (b200vf_ctx *ctx, const uint8_t id[128], int rank, int nranks, b200vf_comm **out);
void b200vf_comm_destroy (b200vf_comm *comm);
/* Even split of `height` rows into nranks blocks with even row0. */
int b200vf_shard_rows (int height, int rank, int nranks, int *row0, int *rows);
/* Halo exchange for a row-sharded buffer of nframes frames: each rank's shard
 * of `rows` rows (row_bytes each) sits at d_shard + halo*row_bytes inside a
 * buffer with `halo` rows of head-room above and below per frame
 * (frame_stride bytes between frames); sends its first/last `halo` rows to the
 * upper/lower neighbour and receives theirs into the head-room, one grouped
 * ncclSend/ncclRecv per neighbour on `stream`. */
int b200vf_comm_halo_exchange (b200vf_comm *comm, uint8_t *d_buf, size_t row_bytes, int rows, int halo,
    size_t frame_stride, int nframes, void *stream);
/* The same exchange in two phases, on the communicator's own stream: begin orders it after the work queued on `stream`
 * so far, end makes `stream` wait for it. In between the caller launches what does not read the halo rows (the
 * interior rows of every shard), so pack / NCCL / unpack overlap the kernel instead of preceding it. */
int b200vf_comm_halo_begin (b200vf_comm *comm, uint8_t *d_buf, size_t row_bytes, int rows, int halo,
    size_t frame_stride, int nframes, void *stream);
int b200vf_comm_halo_end (b200vf_comm *comm, void *stream);
int b200vf_comm_barrier (b200vf_comm *comm, void *stream);
/* All-gather of row shards (geometrictransform: the gather may read any source row): every rank
 * holds a full-size frame buffer with its own rows [row0,row0+rows) (b200vf_shard_rows) filled in; afterwards
 * all `full_rows` rows are valid on every rank. One grouped ncclSend/ncclRecv set per frame. */
int b200vf_comm_allgather_rows (b200vf_comm *comm, uint8_t *d_full, size_t row_bytes, int full_rows,
    size_t frame_stride, int nframes, void *stream);
/* Banded variant: smooth maps (fisheye, bulge, ...) read a bounded band of source rows per output shard.
 * need_lo/need_hi[r] = the source rows [lo,hi) rank r's output shard reads (b200vf_gt_index_row_range over its
 * part of the index table; every rank builds the same table, so every rank knows everybody's band). Each rank
 * sends the part of its own rows a peer needs and receives the part of its band it does not own. */
int b200vf_comm_exchange_rows (b200vf_comm *comm, uint8_t *d_full, size_t row_bytes, int full_rows,
    const int *need_lo, const int *need_hi, size_t frame_stride, int nframes, void *stream);
/* min / max+1 source row referenced by index[0..n) (entries < 0 ignored); lo = hi = 0 when none. */
int b200vf_gt_index_row_range (const int32_t *index, size_t n, int width, int *row_lo, int *row_hi);

/* ----------------------------------------------------- element mirror (host)
 * A GLib-free mirror of the reference's element surface so pipelines can be
 * driven, and parity tests written, the way the reference's would be:
 * factory name -> element, GObject-style properties by name with the
 * reference's ranges/defaults, caps-style negotiation, and the
 * transform vfunc on HOST buffers (upload -> kernel -> download through the
 * HBM pool, the path a sysmem pipeline takes) or on device buffers.
 * Factory names: bayer2rgb rgb2bayer burn chromium dilate dodge exclusion
 * gaussianblur solarize coloreffects chromahold + the geometrictransform set. */
int b200vf_element_factory_make (b200vf_ctx *ctx, const char *factory, b200vf_element **out);
void b200vf_element_destroy (b200vf_element *e);
const char *b200vf_element_factory_name (const b200vf_element *e);
/* Properties as doubles (uint/int/bool/enum/double all fit); enum properties
 * also by nick through set_property_string ("preset"="sepia", "off-edge-pixels"="clamp"). */
int b200vf_element_set_property (b200vf_element *e, const char *name, double value);
int b200vf_element_set_property_string (b200vf_element *e, const char *name, const char *value);
int b200vf_element_get_property (const b200vf_element *e, const char *name, double *value);
/* Negotiation: format strings as in caps ("bggr", "RGBA", "AYUV", ...).
 * Returns B200VF_E_UNSUPPORTED for formats outside the element's pad template. */
int b200vf_element_set_caps (b200vf_element *e, const char *in_format, const char *out_format,
    int width, int height);
/* Frame sizes implied by the negotiated caps (get_unit_size, gstbayer2rgb.c:324-352). */
int b200vf_element_unit_size (const b200vf_element *e, size_t *in_bytes, size_t *out_bytes);
/* transform / transform_frame / transform_frame_ip on host memory: nframes
 * frames packed back to back; pipelined H2D / kernel / D2H; synchronous. */
int b200vf_element_transform_host (b200vf_element *e, const void *h_in, void *h_out, int nframes);
/* Frame layouts. transform_host assumes the default GstVideoInfo layout of the negotiated caps; a GstVideoFrame may
 * carry other plane strides / offsets (GstVideoMeta), which the reference elements honour through
 * GST_VIDEO_FRAME_PLANE_STRIDE / _PLANE_DATA (gstcoloreffects.c:315-329, gstgeometrictransform.c:226-293).
 * b200vf_element_default_layout reports what transform_host assumes on the sink (side 0) / src (side 1) side;
 * b200vf_element_transform_host_layout transforms ONE frame whose planes lie at `offset[i]` from h_in / h_out with
 * `stride[i]` bytes per row (NULL layout = the default one; row_bytes / rows of the argument are ignored). In-place
 * elements pass h_in == h_out and the same layout twice. */
typedef struct b200vf_frame_layout {
  int n_planes;
  size_t offset[4];
  int stride[4];
  int row_bytes[4];             /* bytes of a row that carry samples */
  int rows[4];
} b200vf_frame_layout;
int b200vf_element_default_layout (const b200vf_element *e, int side, b200vf_frame_layout *out);
int b200vf_element_transform_host_layout (b200vf_element *e, const void *h_in, const b200vf_frame_layout *in_layout,
    void *h_out, const b200vf_frame_layout *out_layout);
/* How transform_host treats the caller's host buffers: 0 (default) as they are - pinned buffers copy asynchronously,
 * pageable ones through the driver's staging; 1 = pageable buffers that recur (what a sysmem GstBufferPool hands out):
 * each range is page-locked in place on first sight (cudaHostRegister) and remembered in an LRU cache of 64 ranges. */
int b200vf_element_set_host_mode (b200vf_element *e, int mode);
/* Unlock and forget every range mode 1 page-locked (call before freeing such buffers). */
int b200vf_host_pin_cache_clear (void);
/* The same vfunc on device memory, asynchronous on `stream`. For in-place
 * elements d_out may equal d_in. */
int b200vf_element_transform_device (b200vf_element *e, const void *d_in, void *d_out, int nframes, void *stream);
/* The same vfunc on memories (the GstBuffer path of the HBM pool): frames stay in HBM between elements; per-pixel
 * elements that can join a pending chain only record themselves (see b200vf_memory). in == out for the
 * transform_frame_ip elements. nframes frames packed back to back in each memory. */
int b200vf_element_transform (b200vf_element *e, b200vf_memory *in, b200vf_memory *out, int nframes, void *stream);
/* scenechange only: flags[i] = 1 when frame i of the LAST transform call is a
 * scene change, i.e. where the reference pushes its downstream force-key-unit
 * event (gstscenechange.c:246-257). Returns the number of frames of that call. */
int b200vf_element_last_events (const b200vf_element *e, int *flags, int capacity);
/* videoanalyse: values = (luma-average, luma-variance); simplevideomarkdetect: (message posted 0/1, have-pattern 0/1,
 * data) - of the last frame transformed: the fields of the element messages the reference posts
 * (gstvideoanalyse.c:178-204, gstsimplevideomarkdetect.c:352-389). Returns how many values there are. */
int b200vf_element_last_values (const b200vf_element *e, double *values, int capacity);
/* diffuse only: the seed of its draws and the frame number the next frame draws with (defaults: a fixed seed, 0) */
int b200vf_element_set_rng_seed (b200vf_element *e, uint64_t seed, uint64_t next_frame);
int b200vf_element_get_rng_state (b200vf_element *e, uint64_t *seed, uint64_t *next_frame);

/* ------------------------------------------------- factory introspection
 * What gst-inspect prints for each element, so that the C/GLib shells (gst/gstb200vf.c) can register the
 * factories, install the GObject properties and build the pad templates from one table instead of
 * repeating it (GST_ELEMENT_REGISTER + class_init of each reference element, e.g. gst/gaudieffects/gstburn.c:
 * 120-160; golden dump docs/plugins/gst_plugins_cache.json). */
typedef struct b200vf_factory_info {
  const char *factory;            /* "burn" */
  const char *plugin;             /* "gaudieffects" */
  const char *plugin_description;
  const char *plugin_license;
  const char *type_name;          /* "GstBurn" */
  const char *parent_type_name;   /* "GstVideoFilter", "GstBaseTransform", "GstGeometricTransform", "GstCircleGeometricTransform" */
  const char *klass;              /* "Filter/Effect/Video" */
  const char *long_name;
  const char *description;
  const char *author;
  int in_place;                   /* 1: transform_frame_ip (coloreffects, chromahold) */
  int n_properties;
  int n_formats;                  /* raw video formats of the pad template (sink == src except the bayer elements) */
} b200vf_factory_info;
typedef enum b200vf_prop_type { B200VF_PROP_UINT = 0, B200VF_PROP_INT, B200VF_PROP_BOOL, B200VF_PROP_DOUBLE, B200VF_PROP_ENUM,
  B200VF_PROP_UINT64 /* carried as a double: exact up to 2^53 */ } b200vf_prop_type;
typedef struct b200vf_property_info {
  const char *name;
  int type;                       /* b200vf_prop_type */
  double min, max, def;
  int controllable;               /* GST_PARAM_CONTROLLABLE */
  int n_nicks;                    /* enum nicks, value i = nicks[i] */
  const char *const *nicks;
} b200vf_property_info;
int b200vf_factory_count (void);
int b200vf_factory_get (int index, b200vf_factory_info *out);
int b200vf_factory_find (const char *factory, b200vf_factory_info *out);
int b200vf_factory_property (const char *factory, int index, b200vf_property_info *out);
const char *b200vf_factory_format (const char *factory, int index);

#ifdef __cplusplus
}
#endif
#endif /* B200VF_H */
