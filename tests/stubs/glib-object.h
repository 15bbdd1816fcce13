/* stub: the part of <glib-object.h> the shells use (see README.md) */
#ifndef STUB_GLIB_OBJECT_H
#define STUB_GLIB_OBJECT_H
#include <glib.h>
typedef gsize GType;
typedef struct _GTypeClass { GType g_type; } GTypeClass;
typedef struct _GTypeInstance { GTypeClass *g_class; } GTypeInstance;
typedef struct _GValue { GType g_type; union { gint v_int; gdouble v_double; gpointer v_pointer; } data[2]; } GValue;
#define G_VALUE_INIT { 0, { { 0 } } }
typedef struct _GValueArray { guint n_values; GValue *values; guint n_prealloced; } GValueArray;
typedef struct _GParamSpec { GTypeInstance g_type_instance; const gchar *name; guint flags; GType value_type; GType owner_type; } GParamSpec;
typedef enum { G_PARAM_READABLE = 1 << 0, G_PARAM_WRITABLE = 1 << 1, G_PARAM_READWRITE = 3, G_PARAM_STATIC_NAME = 1 << 5,
  G_PARAM_STATIC_NICK = 1 << 6, G_PARAM_STATIC_BLURB = 1 << 7, G_PARAM_STATIC_STRINGS = (1 << 5) | (1 << 6) | (1 << 7) } GParamFlags;
typedef struct _GObject { GTypeInstance g_type_instance; guint ref_count; gpointer qdata; } GObject;
typedef struct _GObjectClass GObjectClass;
struct _GObjectClass {
  GTypeClass g_type_class;
  void (*set_property) (GObject * object, guint property_id, const GValue * value, GParamSpec * pspec);
  void (*get_property) (GObject * object, guint property_id, GValue * value, GParamSpec * pspec);
  void (*dispose) (GObject * object);
  void (*finalize) (GObject * object);
};
typedef struct _GEnumValue { gint value; const gchar *value_name; const gchar *value_nick; } GEnumValue;
typedef void (*GBaseInitFunc) (gpointer g_class);
typedef void (*GBaseFinalizeFunc) (gpointer g_class);
typedef void (*GClassInitFunc) (gpointer g_class, gpointer class_data);
typedef void (*GClassFinalizeFunc) (gpointer g_class, gpointer class_data);
typedef void (*GInstanceInitFunc) (GTypeInstance * instance, gpointer g_class);
typedef struct _GTypeInfo {
  guint16 class_size; GBaseInitFunc base_init; GBaseFinalizeFunc base_finalize; GClassInitFunc class_init;
  GClassFinalizeFunc class_finalize; gconstpointer class_data; guint16 instance_size; guint16 n_preallocs;
  GInstanceInitFunc instance_init; const gpointer value_table;
} GTypeInfo;
typedef enum { G_TYPE_FLAG_NONE = 0, G_TYPE_FLAG_ABSTRACT = 1 << 4 } GTypeFlags;
#define G_TYPE_FUNDAMENTAL_SHIFT 2
#define G_TYPE_MAKE_FUNDAMENTAL(x) ((GType) ((x) << G_TYPE_FUNDAMENTAL_SHIFT))
#define G_TYPE_BOOLEAN G_TYPE_MAKE_FUNDAMENTAL (5)
#define G_TYPE_INT G_TYPE_MAKE_FUNDAMENTAL (6)
#define G_TYPE_UINT G_TYPE_MAKE_FUNDAMENTAL (7)
#define G_TYPE_UINT64 G_TYPE_MAKE_FUNDAMENTAL (11)
#define G_TYPE_ENUM G_TYPE_MAKE_FUNDAMENTAL (12)
#define G_TYPE_DOUBLE G_TYPE_MAKE_FUNDAMENTAL (15)
#define G_TYPE_OBJECT G_TYPE_MAKE_FUNDAMENTAL (20)
gboolean g_type_check_value_holds (const GValue * value, GType type);
#define G_VALUE_HOLDS(value, type) (g_type_check_value_holds ((value), (type)))
#define G_VALUE_HOLDS_BOOLEAN(value) (G_VALUE_HOLDS ((value), G_TYPE_BOOLEAN))
#define G_VALUE_HOLDS_INT(value) (G_VALUE_HOLDS ((value), G_TYPE_INT))
#define G_VALUE_HOLDS_UINT(value) (G_VALUE_HOLDS ((value), G_TYPE_UINT))
#define G_VALUE_HOLDS_UINT64(value) (G_VALUE_HOLDS ((value), G_TYPE_UINT64))
#define G_VALUE_HOLDS_DOUBLE(value) (G_VALUE_HOLDS ((value), G_TYPE_DOUBLE))
#define G_VALUE_HOLDS_ENUM(value) (G_VALUE_HOLDS ((value), G_TYPE_ENUM))
#define G_OBJECT_CLASS(klass) ((GObjectClass *) (klass))
#define G_OBJECT_GET_CLASS(object) ((GObjectClass *) (((GTypeInstance *) (object))->g_class))
#define G_OBJECT_TYPE(object) (((GTypeInstance *) (object))->g_class->g_type)
GValue *g_value_init (GValue * value, GType g_type);
void g_value_unset (GValue * value);
guint g_value_get_uint (const GValue * value);
guint64 g_value_get_uint64 (const GValue * value);
void g_value_set_uint64 (GValue * value, guint64 v);
gint g_value_get_int (const GValue * value);
gboolean g_value_get_boolean (const GValue * value);
gdouble g_value_get_double (const GValue * value);
gint g_value_get_enum (const GValue * value);
gpointer g_value_get_boxed (const GValue * value);
void g_value_set_uint (GValue * value, guint v);
void g_value_set_int (GValue * value, gint v);
void g_value_set_boolean (GValue * value, gboolean v);
void g_value_set_double (GValue * value, gdouble v);
void g_value_set_enum (GValue * value, gint v);
void g_value_take_boxed (GValue * value, gconstpointer v_boxed);
GValueArray *g_value_array_new (guint n_prealloced);
GValueArray *g_value_array_append (GValueArray * value_array, const GValue * value);
GValue *g_value_array_get_nth (GValueArray * value_array, guint index_);
GParamSpec *g_param_spec_uint (const gchar * name, const gchar * nick, const gchar * blurb, guint minimum, guint maximum, guint default_value, GParamFlags flags);
GParamSpec *g_param_spec_uint64 (const gchar * name, const gchar * nick, const gchar * blurb, guint64 minimum, guint64 maximum, guint64 default_value, GParamFlags flags);
GParamSpec *g_param_spec_int (const gchar * name, const gchar * nick, const gchar * blurb, gint minimum, gint maximum, gint default_value, GParamFlags flags);
GParamSpec *g_param_spec_boolean (const gchar * name, const gchar * nick, const gchar * blurb, gboolean default_value, GParamFlags flags);
GParamSpec *g_param_spec_double (const gchar * name, const gchar * nick, const gchar * blurb, gdouble minimum, gdouble maximum, gdouble default_value, GParamFlags flags);
GParamSpec *g_param_spec_enum (const gchar * name, const gchar * nick, const gchar * blurb, GType enum_type, gint default_value, GParamFlags flags);
GParamSpec *g_param_spec_value_array (const gchar * name, const gchar * nick, const gchar * blurb, GParamSpec * element_spec, GParamFlags flags);
void g_object_class_install_property (GObjectClass * oclass, guint property_id, GParamSpec * pspec);
gpointer g_object_new (GType object_type, const gchar * first_property_name, ...);
GType g_type_from_name (const gchar * name);
gboolean g_type_is_a (GType type, GType is_a_type);
GType g_type_register_static (GType parent_type, const gchar * type_name, const GTypeInfo * info, GTypeFlags flags);
GType g_type_register_static_simple (GType parent_type, const gchar * type_name, guint class_size, GClassInitFunc class_init,
    guint instance_size, GInstanceInitFunc instance_init, GTypeFlags flags);
gpointer g_type_class_peek_parent (gpointer g_class);
GType g_enum_register_static (const gchar * name, const GEnumValue * const_static_values);
/* G_DEFINE_TYPE: prototypes of <prefix>_class_init / _init, <prefix>_parent_class, <prefix>_get_type () */
#define G_DEFINE_TYPE(TypeName, type_name, TYPE_PARENT) \
  static void type_name##_init (TypeName * self); \
  static void type_name##_class_init (TypeName##Class * klass); \
  static gpointer type_name##_parent_class = NULL; \
  static void type_name##_class_intern_init (gpointer klass, gpointer data) { \
    type_name##_parent_class = g_type_class_peek_parent (klass); \
    type_name##_class_init ((TypeName##Class *) klass); \
  } \
  static void type_name##_instance_intern_init (GTypeInstance * inst, gpointer klass) { type_name##_init ((TypeName *) inst); } \
  GType type_name##_get_type (void) { \
    static GType t = 0; \
    if (!t) t = g_type_register_static_simple (TYPE_PARENT, #TypeName, sizeof (TypeName##Class), type_name##_class_intern_init, \
          sizeof (TypeName), type_name##_instance_intern_init, G_TYPE_FLAG_NONE); \
    return t; \
  }
#endif
