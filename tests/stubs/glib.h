/* stub: the part of <glib.h> the shells use (see README.md) */
#ifndef STUB_GLIB_H
#define STUB_GLIB_H
#include <stddef.h>
#include <stdarg.h>
typedef char gchar;
typedef int gint;
typedef unsigned int guint;
typedef int gboolean;
typedef unsigned char guint8;
typedef unsigned short guint16;
typedef unsigned int guint32;
typedef unsigned long long guint64;
typedef long long gint64;
typedef double gdouble;
typedef float gfloat;
typedef void *gpointer;
typedef const void *gconstpointer;
typedef unsigned long gsize;
typedef long gssize;
typedef unsigned long gulong;
#define TRUE 1
#define FALSE 0
#ifndef NULL
#define NULL ((void *) 0)
#endif
#define G_BEGIN_DECLS
#define G_END_DECLS
#define G_MAXDOUBLE 1.7976931348623157e308
#define G_MAXUINT64 0xffffffffffffffffULL
#define G_GSIZE_FORMAT "lu"
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#define G_UNLIKELY(x) (x)
typedef struct _GString { gchar *str; gsize len; gsize allocated_len; } GString;
gpointer g_malloc0 (gsize n_bytes);
void g_free (gpointer mem);
#define g_new0(type, n) ((type *) g_malloc0 (sizeof (type) * (n)))
gchar *g_strdup_printf (const gchar * format, ...) __attribute__ ((format (printf, 1, 2)));
gint g_snprintf (gchar * string, gulong n, gchar const *format, ...) __attribute__ ((format (printf, 3, 4)));
GString *g_string_new (const gchar * init);
GString *g_string_append (GString * string, const gchar * val);
void g_string_append_printf (GString * string, const gchar * format, ...) __attribute__ ((format (printf, 2, 3)));
gchar *g_string_free (GString * string, gboolean free_segment);
void g_return_if_fail_warning (const char *domain, const char *func, const char *expr);
#define g_return_val_if_fail(expr, val) do { if (!(expr)) { g_return_if_fail_warning (NULL, __func__, #expr); return (val); } } while (0)
#define g_return_val_if_reached(val) do { g_return_if_fail_warning (NULL, __func__, "reached"); return (val); } while (0)
#endif
