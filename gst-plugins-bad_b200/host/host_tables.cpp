// host_tables.cpp - host-side constant builders that must stay on the CPU for
// bit-exactness: coloreffects preset tables and the gaussian
// taps (libm pow/sqrt), exactly as the reference computes them.
#include "../csrc/common.cuh"
#include <math.h>
#include "coloreffects_tables.inc"

// preset numbering = GstColorEffectsPreset (gst/coloreffects/gstcoloreffects.c:74-96);
// (table, map_luma) pairs as in set_property (:503-548).
B200VF_API int b200vf_coloreffects_table (int preset, const uint8_t **table768, int *map_luma) {
  B200VF_REQUIRE (table768 && map_luma, B200VF_E_INVAL, "coloreffects_table: NULL argument");
  switch (preset) {
    case 1: *table768 = k_heat_table; *map_luma = 1; return B200VF_OK;
    case 2: *table768 = k_sepia_table; *map_luma = 1; return B200VF_OK;
    case 3: *table768 = k_xray_table; *map_luma = 1; return B200VF_OK;
    case 4: *table768 = k_xpro_table; *map_luma = 0; return B200VF_OK;
    case 5: *table768 = k_yellowblue_table; *map_luma = 0; return B200VF_OK;
    case 0: *table768 = nullptr; *map_luma = 0; return B200VF_OK;     // none: the element is a no-op
    default:
      b200vf_set_error ("coloreffects_table: preset %d not in [0,5]", preset);
      return B200VF_E_PROPERTY;
  }
}

// make_gaussian_kernel, gst/gaudieffects/gstgaussblur.c:361-422, with the same
// float/double mix: fe and dx are floats computed in double, each tap is
// dx*pow(e, fe*i*i) rounded to float, the running sums are float.
B200VF_API int b200vf_gauss_kernel (float sigma, float *kernel, float *kernel_sum, int capacity) {
  B200VF_REQUIRE (kernel && kernel_sum, B200VF_E_INVAL, "gauss_kernel: NULL argument");
  B200VF_REQUIRE (sigma >= -20.0f && sigma <= 20.0f, B200VF_E_PROPERTY, "gauss_kernel: sigma %g not in [-20,20]", sigma);
  const double kPi = 3.1415926535897932384626433832795028841971693993751;   // G_PI
  const double kE = 2.7182818284590452353602874713526624977572470937000;    // G_E
  const float fe = -0.5 / (sigma * sigma);
  const float dx = 1.0 / (sigma * sqrt (2 * kPi));
  int center = (int) ceil (2.5 * fabs (sigma));
  int ws = 1 + 2 * center;
  B200VF_REQUIRE (ws <= capacity, B200VF_E_INVAL, "gauss_kernel: window %d exceeds capacity %d", ws, capacity);
  if (ws == 1) {
    kernel[0] = 1.0f;
    kernel_sum[0] = 1.0f;
    return 1;
  }
  float sum = kernel[center] = dx;
  for (int i = 1; i <= center; i++) {
    float fx = dx * pow (kE, fe * i * i);
    kernel[center + i] = kernel[center - i] = fx;
    sum += 2 * fx;
  }
  if (sigma < 0) {                 // a negative sigma sharpens (:395-398)
    sum = -sum;
    kernel[center] += 2.0 * sum;
  }
  for (int i = 0; i < ws; i++) kernel[i] /= sum;
  float acc = 0.0f;
  for (int i = 0; i < ws; i++) {
    acc += kernel[i];
    kernel_sum[i] = acc;
  }
  return ws;
}

// Halo rows a row shard of b200vf_gaussblur must be given above and below (unless at the frame's edge):
// `center` rows of the window, plus one when the blurred channels start p0 > 0 bytes into the pixel and rows are
// unpadded - the last p0 bytes of a row's last pixel are then the first bytes of the next row, so the
// window's outermost rows reach one row further. b200vf_comm_halo_exchange moves exactly these rows.
B200VF_API int b200vf_gaussblur_halo_rows (int windowsize, int p0, int stride, int width) {
  B200VF_REQUIRE (windowsize >= 1 && (windowsize & 1) && p0 >= 0 && p0 <= 3 && width > 0 && stride >= 4 * width, B200VF_E_INVAL,
      "gaussblur_halo_rows: bad argument");
  return windowsize / 2 + ((p0 > 0 && stride == 4 * width && windowsize > 1) ? 1 : 0);
}
