/* napi_mock.c — a miniature Node-API runtime for tests: just enough of the stable C ABI (exactly the functions
 * homography.js_b200/js/napi_min.h declares) to EXECUTE the addon js/hgwarp_napi.c in an image that has no Node.js.
 *
 * It is linked with the addon into one shared library (tests/napi_mock.py builds it); the addon's NAPI_MODULE
 * constructor calls napi_module_register() below, mock_env_create() runs the module's Init and keeps the exports
 * object, and mock_call() invokes an exported function the way Node would: a napi_callback_info carrying argc / argv.
 * Values are plain tagged structs owned by the environment (freed, finalizers run, in mock_env_destroy).
 * TEST INFRASTRUCTURE ONLY: nothing in the product links or loads this file. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../homography.js_b200/js/napi_min.h"

typedef enum { V_UNDEFINED = 0, V_NUMBER = 1, V_EXTERNAL = 2, V_OBJECT = 3, V_ARRAYBUFFER = 4, V_TYPEDARRAY = 5, V_FUNCTION = 6 } vkind;

#define MAX_PROPS 32
struct prop {
    char name[40];
    struct napi_value__ *value;
};
struct napi_value__ {
    vkind kind;
    double num;
    /* external */
    void *ext;
    napi_finalize fin;
    void *fin_hint;
    /* arraybuffer */
    void *data;
    size_t bytes;
    int owned;
    /* typedarray */
    napi_typedarray_type ttype;
    size_t length, byte_offset;
    struct napi_value__ *buffer;
    /* function */
    napi_callback fn;
    /* object */
    struct prop props[MAX_PROPS];
    int nprops;
};
struct napi_env__ {
    struct napi_value__ **all;
    size_t n, cap;
    char exc[800];
    int has_exc;
    struct napi_value__ *exports;
};
struct napi_callback_info__ {
    size_t argc;
    napi_value *argv;
};

static napi_module *g_module = NULL;
void napi_module_register(napi_module *mod) { g_module = mod; }

static size_t elem_size(napi_typedarray_type t)
{
    switch (t) {
        case napi_int8_array: case napi_uint8_array: case napi_uint8_clamped_array: return 1;
        case napi_int16_array: case napi_uint16_array: return 2;
        case napi_int32_array: case napi_uint32_array: case napi_float32_array: return 4;
        default: return 8;
    }
}

static struct napi_value__ *new_value(napi_env env, vkind k)
{
    struct napi_value__ *v = (struct napi_value__ *)calloc(1, sizeof *v);
    if (!v) abort();
    v->kind = k;
    if (env->n == env->cap) {
        env->cap = env->cap ? 2 * env->cap : 64;
        env->all = (struct napi_value__ **)realloc(env->all, env->cap * sizeof *env->all);
        if (!env->all) abort();
    }
    env->all[env->n++] = v;
    return v;
}

/* ------------------------------------------------------------------ the Node-API subset */
napi_status napi_get_cb_info(napi_env env, napi_callback_info info, size_t *argc, napi_value *argv, napi_value *this_arg, void **data)
{
    (void)env;
    if (argc) {
        const size_t room = *argc; /* in: capacity of argv; out: the actual count (like Node) */
        if (argv) {
            for (size_t i = 0; i < room; ++i) {
                if (i < info->argc) argv[i] = info->argv[i];
                else argv[i] = NULL; /* Node fills with undefined; the addon never reads past argc */
            }
        }
        *argc = info->argc;
    }
    if (this_arg) *this_arg = NULL;
    if (data) *data = NULL;
    return napi_ok;
}
napi_status napi_get_value_int32(napi_env env, napi_value v, int32_t *r)
{
    (void)env;
    if (!v || v->kind != V_NUMBER) return napi_number_expected;
    /* ToInt32: NaN / Inf -> 0, otherwise truncate, modulo 2^32 */
    const double d = v->num;
    if (d != d || d - d != 0.0) { *r = 0; return napi_ok; }
    double m = fmod(trunc(d), 4294967296.0);
    if (m < 0) m += 4294967296.0;
    *r = (int32_t)(uint32_t)m;
    return napi_ok;
}
napi_status napi_get_value_double(napi_env env, napi_value v, double *r)
{
    (void)env;
    if (!v || v->kind != V_NUMBER) return napi_number_expected;
    *r = v->num;
    return napi_ok;
}
napi_status napi_get_value_external(napi_env env, napi_value v, void **r)
{
    (void)env;
    if (!v || v->kind != V_EXTERNAL) return napi_invalid_arg;
    *r = v->ext;
    return napi_ok;
}
napi_status napi_create_external(napi_env env, void *data, napi_finalize fin, void *hint, napi_value *r)
{
    struct napi_value__ *v = new_value(env, V_EXTERNAL);
    v->ext = data;
    v->fin = fin;
    v->fin_hint = hint;
    *r = v;
    return napi_ok;
}
napi_status napi_create_int32(napi_env env, int32_t x, napi_value *r)
{
    struct napi_value__ *v = new_value(env, V_NUMBER);
    v->num = (double)x;
    *r = v;
    return napi_ok;
}
napi_status napi_create_double(napi_env env, double x, napi_value *r)
{
    struct napi_value__ *v = new_value(env, V_NUMBER);
    v->num = x;
    *r = v;
    return napi_ok;
}
napi_status napi_get_undefined(napi_env env, napi_value *r)
{
    *r = new_value(env, V_UNDEFINED);
    return napi_ok;
}
napi_status napi_create_object(napi_env env, napi_value *r)
{
    *r = new_value(env, V_OBJECT);
    return napi_ok;
}
napi_status napi_set_named_property(napi_env env, napi_value obj, const char *name, napi_value value)
{
    (void)env;
    if (!obj || obj->kind != V_OBJECT) return napi_object_expected;
    for (int i = 0; i < obj->nprops; ++i)
        if (!strcmp(obj->props[i].name, name)) { obj->props[i].value = value; return napi_ok; }
    if (obj->nprops == MAX_PROPS || strlen(name) >= sizeof obj->props[0].name) return napi_generic_failure;
    strcpy(obj->props[obj->nprops].name, name);
    obj->props[obj->nprops++].value = value;
    return napi_ok;
}
napi_status napi_get_typedarray_info(napi_env env, napi_value v, napi_typedarray_type *type, size_t *length, void **data,
                                     napi_value *arraybuffer, size_t *byte_offset)
{
    (void)env;
    if (!v || v->kind != V_TYPEDARRAY) return napi_invalid_arg;
    if (type) *type = v->ttype;
    if (length) *length = v->length;
    if (data) *data = (char *)v->buffer->data + v->byte_offset;
    if (arraybuffer) *arraybuffer = v->buffer;
    if (byte_offset) *byte_offset = v->byte_offset;
    return napi_ok;
}
napi_status napi_create_arraybuffer(napi_env env, size_t bytes, void **data, napi_value *r)
{
    struct napi_value__ *v = new_value(env, V_ARRAYBUFFER);
    v->data = calloc(bytes ? bytes : 1, 1); /* Node zero-fills new ArrayBuffers */
    if (!v->data) return napi_generic_failure;
    v->bytes = bytes;
    v->owned = 1;
    if (data) *data = v->data;
    *r = v;
    return napi_ok;
}
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value ab, size_t byte_offset, napi_value *r)
{
    if (!ab || ab->kind != V_ARRAYBUFFER) return napi_invalid_arg;
    const size_t es = elem_size(type);
    if (byte_offset % es != 0 || byte_offset + length * es > ab->bytes) {
        snprintf(env->exc, sizeof env->exc, "RangeError: Invalid typed array length");
        env->has_exc = 1;
        return napi_pending_exception;
    }
    struct napi_value__ *v = new_value(env, V_TYPEDARRAY);
    v->ttype = type;
    v->length = length;
    v->byte_offset = byte_offset;
    v->buffer = ab;
    *r = v;
    return napi_ok;
}
napi_status napi_throw_error(napi_env env, const char *code, const char *msg)
{
    snprintf(env->exc, sizeof env->exc, "%s%s%s", code ? code : "", code ? ": " : "", msg ? msg : "");
    env->has_exc = 1;
    return napi_ok;
}
napi_status napi_define_properties(napi_env env, napi_value obj, size_t n, const napi_property_descriptor *p)
{
    for (size_t i = 0; i < n; ++i) {
        napi_value v = p[i].value;
        if (p[i].method) {
            v = new_value(env, V_FUNCTION);
            v->fn = p[i].method;
        }
        const napi_status st = napi_set_named_property(env, obj, p[i].utf8name, v);
        if (st != napi_ok) return st;
    }
    return napi_ok;
}

/* ------------------------------------------------------------------ driver API (ctypes) */
#define API __attribute__((visibility("default")))

API void *mock_env_create(void)
{
    if (!g_module || !g_module->nm_register_func) return NULL;
    napi_env env = (napi_env)calloc(1, sizeof *env);
    if (!env) return NULL;
    env->exports = new_value(env, V_OBJECT);
    napi_value r = g_module->nm_register_func(env, env->exports);
    if (r && r->kind == V_OBJECT) env->exports = r;
    return env;
}
API void mock_env_destroy(void *e)
{
    napi_env env = (napi_env)e;
    if (!env) return;
    for (size_t i = 0; i < env->n; ++i) {
        struct napi_value__ *v = env->all[i];
        if (v->kind == V_EXTERNAL && v->fin) v->fin(env, v->ext, v->fin_hint); /* what the garbage collector would do */
        if (v->kind == V_ARRAYBUFFER && v->owned) free(v->data);
        free(v);
    }
    free(env->all);
    free(env);
}
API const char *mock_module_name(void) { return g_module ? g_module->nm_modname : NULL; }
API int mock_export_count(void *e) { return ((napi_env)e)->exports->nprops; }
API const char *mock_export_name(void *e, int i) { return ((napi_env)e)->exports->props[i].name; }
API void *mock_number(void *e, double x)
{
    napi_value v;
    napi_create_double((napi_env)e, x, &v);
    return v;
}
API void *mock_undefined(void *e)
{
    napi_value v;
    napi_get_undefined((napi_env)e, &v);
    return v;
}
/* a typed array over CALLER memory (a numpy buffer): borrowed, like a JS typed array handed to a native call */
API void *mock_typedarray(void *e, int type, void *data, size_t length)
{
    napi_env env = (napi_env)e;
    struct napi_value__ *ab = new_value(env, V_ARRAYBUFFER);
    ab->data = data;
    ab->bytes = length * elem_size((napi_typedarray_type)type);
    ab->owned = 0;
    struct napi_value__ *v = new_value(env, V_TYPEDARRAY);
    v->ttype = (napi_typedarray_type)type;
    v->length = length;
    v->buffer = ab;
    return v;
}
/* exports.<name>(argv...) -> value, or NULL with a pending exception */
API void *mock_call(void *e, const char *name, int argc, void **argv)
{
    napi_env env = (napi_env)e;
    env->has_exc = 0;
    env->exc[0] = 0;
    for (int i = 0; i < env->exports->nprops; ++i) {
        struct prop *p = &env->exports->props[i];
        if (!strcmp(p->name, name) && p->value && p->value->kind == V_FUNCTION) {
            struct napi_callback_info__ info = {(size_t)argc, (napi_value *)argv};
            napi_value r = p->value->fn(env, &info);
            if (env->has_exc) return NULL;
            return r ? (void *)r : mock_undefined(env); /* a NULL return without an exception is `undefined` */
        }
    }
    snprintf(env->exc, sizeof env->exc, "TypeError: native.%s is not a function", name);
    env->has_exc = 1;
    return NULL;
}
API const char *mock_exception(void *e) { return ((napi_env)e)->has_exc ? ((napi_env)e)->exc : NULL; }
API int mock_kind(void *v) { return (int)((napi_value)v)->kind; }
API double mock_get_number(void *v) { return ((napi_value)v)->num; }
API int mock_typedarray_get(void *v, int *type, size_t *length, void **data)
{
    napi_typedarray_type t;
    if (napi_get_typedarray_info(NULL, (napi_value)v, &t, length, data, NULL, NULL) != napi_ok) return 1;
    *type = (int)t;
    return 0;
}
API void *mock_get_property(void *v, const char *name)
{
    napi_value o = (napi_value)v;
    if (!o || o->kind != V_OBJECT) return NULL;
    for (int i = 0; i < o->nprops; ++i)
        if (!strcmp(o->props[i].name, name)) return o->props[i].value;
    return NULL;
}
