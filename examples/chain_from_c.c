/* chain_from_c.c - the C-ABI used from plain C, with no Python, torch or GLib in the process: what a shell
 * (gst/gstb200vf.c) does for `bayer2rgb ! coloreffects preset=sepia ! solarize`, reduced to its calls.
 *
 *   gcc -std=c99 -Iinclude examples/chain_from_c.c -Lgst-plugins-bad_b200/lib -lb200vf -Wl,-rpath,... -o chain_from_c
 *   ./chain_from_c 1920 1080
 *
 * Prints one line: width height fnv1a64(memories path) fnv1a64(host-buffer path) launches h2d d2h - the two hashes
 * must agree (same bytes whether the frame stays in HBM between the elements or crosses PCIe for each), and
 * tests/test_c_example.py compares them with the oracle's chain on the same input. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200vf.h"

#define CHECK(call)                                                                   \
  do {                                                                                \
    int rc_ = (call);                                                                 \
    if (rc_ != B200VF_OK) {                                                           \
      fprintf (stderr, "%s: %s (%s)\n", #call, b200vf_status_string (rc_), b200vf_last_error ()); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

static uint64_t fnv1a64 (const uint8_t *p, size_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
  return h;
}

/* the input pattern: a 32-bit LCG, top byte of each state (tests regenerate it in numpy) */
static void fill_pattern (uint8_t *p, size_t n) {
  uint32_t s = 12345u;
  for (size_t i = 0; i < n; i++) { s = s * 1664525u + 1013904223u; p[i] = (uint8_t) (s >> 24); }
}

int main (int argc, char **argv) {
  const int w = argc > 1 ? atoi (argv[1]) : 640, h = argc > 2 ? atoi (argv[2]) : 480;
  const size_t in_bytes = (size_t) w * h, out_bytes = 4 * in_bytes;
  b200vf_ctx *ctx = NULL;
  b200vf_element *bayer = NULL, *color = NULL, *solar = NULL;
  b200vf_memory *m0 = NULL, *m1 = NULL, *m2 = NULL;
  void *p = NULL;

  CHECK (b200vf_ctx_create (0, &ctx));
  CHECK (b200vf_element_factory_make (ctx, "bayer2rgb", &bayer));
  CHECK (b200vf_element_factory_make (ctx, "coloreffects", &color));
  CHECK (b200vf_element_factory_make (ctx, "solarize", &solar));
  CHECK (b200vf_element_set_caps (bayer, "bggr", "BGRx", w, h));
  CHECK (b200vf_element_set_caps (color, "BGRx", "BGRx", w, h));
  CHECK (b200vf_element_set_caps (solar, "BGRx", "BGRx", w, h));
  CHECK (b200vf_element_set_property_string (color, "preset", "sepia"));

  /* 1. frames as memories of the HBM pool: the elements record themselves, the host read launches ONE kernel */
  CHECK (b200vf_memory_new (ctx, in_bytes, &m0));
  CHECK (b200vf_memory_new (ctx, out_bytes, &m1));
  CHECK (b200vf_memory_new (ctx, out_bytes, &m2));
  uint64_t c0[4], c1[4];
  CHECK (b200vf_ctx_transfer_counts (ctx, &c0[0], &c0[1], &c0[2], &c0[3]));
  const uint64_t l0 = b200vf_ctx_launch_count (ctx);
  CHECK (b200vf_memory_map (m0, B200VF_MAP_WRITE, &p, NULL));
  fill_pattern ((uint8_t *) p, in_bytes);
  CHECK (b200vf_memory_unmap (m0));
  CHECK (b200vf_element_transform (bayer, m0, m1, 1, NULL));
  CHECK (b200vf_element_transform (color, m1, m1, 1, NULL));      /* transform_frame_ip */
  CHECK (b200vf_element_transform (solar, m1, m2, 1, NULL));
  CHECK (b200vf_memory_map (m2, B200VF_MAP_READ, &p, NULL));
  const uint64_t hash_mem = fnv1a64 ((const uint8_t *) p, out_bytes);
  CHECK (b200vf_memory_unmap (m2));
  const uint64_t launches = b200vf_ctx_launch_count (ctx) - l0;
  CHECK (b200vf_ctx_transfer_counts (ctx, &c1[0], &c1[1], &c1[2], &c1[3]));

  /* 2. the same chain on caller-owned (pageable) host buffers, one element at a time */
  uint8_t *in = (uint8_t *) malloc (in_bytes), *a = (uint8_t *) malloc (out_bytes), *b = (uint8_t *) malloc (out_bytes);
  if (!in || !a || !b) return 2;
  fill_pattern (in, in_bytes);
  CHECK (b200vf_element_transform_host (bayer, in, a, 1));
  CHECK (b200vf_element_transform_host (color, a, a, 1));
  CHECK (b200vf_element_transform_host (solar, a, b, 1));
  const uint64_t hash_host = fnv1a64 (b, out_bytes);

  printf ("%d %d %016llx %016llx %llu %llu %llu\n", w, h, (unsigned long long) hash_mem, (unsigned long long) hash_host,
      (unsigned long long) launches, (unsigned long long) (c1[0] - c0[0]), (unsigned long long) (c1[2] - c0[2]));

  free (in); free (a); free (b);
  b200vf_memory_unref (m0); b200vf_memory_unref (m1); b200vf_memory_unref (m2);
  b200vf_element_destroy (bayer); b200vf_element_destroy (color); b200vf_element_destroy (solar);
  b200vf_ctx_destroy (ctx);
  return 0;
}
