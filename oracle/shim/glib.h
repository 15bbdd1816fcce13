/* Minimal GLib typedef/macro shim used ONLY to compile the reference's own
 * inner loops (ORC *-dist.c C backups and line-range extracted static loops)
 * into oracle/_ref/.  Test infrastructure, never linked into the product. */
#ifndef B200VF_ORACLE_GLIB_SHIM_H
#define B200VF_ORACLE_GLIB_SHIM_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <limits.h>
typedef uint8_t guint8;
typedef int8_t gint8;
typedef uint16_t guint16;
typedef int16_t gint16;
typedef uint32_t guint32;
typedef int32_t gint32;
typedef uint64_t guint64;
typedef int64_t gint64;
typedef int gint;
typedef unsigned int guint;
typedef int gboolean;
typedef float gfloat;
typedef double gdouble;
typedef char gchar;
typedef unsigned char guchar;
typedef size_t gsize;
typedef void *gpointer;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#ifndef MIN
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#endif
#define CLAMP(x, low, high) (((x) > (high)) ? (high) : (((x) < (low)) ? (low) : (x)))
#define ABS(a) (((a) < 0) ? -(a) : (a))
#define G_PI 3.1415926535897932384626433832795028841971693993751
#define G_E 2.7182818284590452353602874713526624977572470937000
#define G_MAXUINT UINT_MAX
#define G_UNLIKELY(x) (x)
#define G_LIKELY(x) (x)
#define g_malloc(n) malloc(n)
#define g_malloc0(n) calloc(1, (n))
#define g_free(p) free(p)
#define g_new(type, n) ((type *) malloc(sizeof(type) * (n)))
#define g_new0(type, n) ((type *) calloc((n), sizeof(type)))
#define g_return_val_if_fail(expr, val) do { if (!(expr)) return (val); } while (0)
#define GST_DEBUG_OBJECT(...) do { } while (0)
#define GST_WARNING_OBJECT(...) do { } while (0)
#define GST_INFO_OBJECT(...) do { } while (0)
#define GST_LOG_OBJECT(...) do { } while (0)
#define GST_DEBUG(...) do { } while (0)
#define GST_ROUND_UP_4(n) (((n) + 3) & ~3)
#endif
