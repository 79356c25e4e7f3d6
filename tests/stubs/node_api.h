/* tests/stubs/node_api.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Node.js (and therefore its node_api.h) is not in the build image.  This stub declares, with the signatures of
 * Node's ABI-stable C N-API (node_api.h / js_native_api.h, N-API version 6), exactly the types, enums and
 * functions binding/jslp_addon.cc uses, so that the addon's translation unit can be compiled and type-checked
 * against include/jslp_b200.h on every CPU test run (tests/test_host_cpu.py::test_napi_addon_compiles).
 * Nothing links against it. */
#ifndef JSLP_TEST_NODE_API_STUB_H
#define JSLP_TEST_NODE_API_STUB_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_ref__ *napi_ref;
typedef struct napi_callback_info__ *napi_callback_info;

typedef enum {
    napi_ok, napi_invalid_arg, napi_object_expected, napi_string_expected, napi_name_expected, napi_function_expected,
    napi_number_expected, napi_boolean_expected, napi_array_expected, napi_generic_failure, napi_pending_exception
} napi_status;
typedef enum {
    napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function,
    napi_external, napi_bigint
} napi_valuetype;
typedef enum {
    napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array,
    napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array
} napi_typedarray_type;
typedef enum {
    napi_default = 0, napi_writable = 1 << 0, napi_enumerable = 1 << 1, napi_configurable = 1 << 2, napi_static = 1 << 10
} napi_property_attributes;

typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void *finalize_data, void *finalize_hint);
typedef struct {
    const char *utf8name;
    napi_value name;
    napi_callback method;
    napi_callback getter;
    napi_callback setter;
    napi_value value;
    napi_property_attributes attributes;
    void *data;
} napi_property_descriptor;

#define NAPI_AUTO_LENGTH SIZE_MAX

napi_status napi_get_undefined(napi_env env, napi_value *result);
napi_status napi_get_null(napi_env env, napi_value *result);
napi_status napi_get_boolean(napi_env env, bool value, napi_value *result);
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype *result);
napi_status napi_coerce_to_bool(napi_env env, napi_value value, napi_value *result);
napi_status napi_get_value_bool(napi_env env, napi_value value, bool *result);
napi_status napi_get_value_double(napi_env env, napi_value value, double *result);
napi_status napi_get_value_string_utf8(napi_env env, napi_value value, char *buf, size_t bufsize, size_t *result);
napi_status napi_create_double(napi_env env, double value, napi_value *result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value *result);
napi_status napi_create_string_utf8(napi_env env, const char *str, size_t length, napi_value *result);
napi_status napi_create_object(napi_env env, napi_value *result);
napi_status napi_create_array_with_length(napi_env env, size_t length, napi_value *result);
napi_status napi_get_array_length(napi_env env, napi_value value, uint32_t *result);
napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value *result);
napi_status napi_set_element(napi_env env, napi_value object, uint32_t index, napi_value value);
napi_status napi_set_named_property(napi_env env, napi_value object, const char *utf8name, napi_value value);
napi_status napi_get_named_property(napi_env env, napi_value object, const char *utf8name, napi_value *result);
napi_status napi_has_named_property(napi_env env, napi_value object, const char *utf8name, bool *result);
napi_status napi_is_typedarray(napi_env env, napi_value value, bool *result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type *type, size_t *length,
                                     void **data, napi_value *arraybuffer, size_t *byte_offset);
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void **data, napi_value *result);
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer,
                                   size_t byte_offset, napi_value *result);
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t *argc, napi_value *argv, napi_value *this_arg,
                             void **data);
napi_status napi_define_class(napi_env env, const char *utf8name, size_t length, napi_callback constructor, void *data,
                              size_t property_count, const napi_property_descriptor *properties, napi_value *result);
napi_status napi_wrap(napi_env env, napi_value js_object, void *native_object, napi_finalize finalize_cb, void *finalize_hint,
                      napi_ref *result);
napi_status napi_unwrap(napi_env env, napi_value js_object, void **result);
napi_status napi_create_reference(napi_env env, napi_value value, uint32_t initial_refcount, napi_ref *result);
napi_status napi_create_function(napi_env env, const char *utf8name, size_t length, napi_callback cb, void *data,
                                 napi_value *result);
napi_status napi_throw_error(napi_env env, const char *code, const char *msg);
napi_status napi_throw_type_error(napi_env env, const char *code, const char *msg);
napi_status napi_throw_range_error(napi_env env, const char *code, const char *msg);

typedef napi_value (*napi_addon_register_func)(napi_env env, napi_value exports);
typedef struct napi_module {
    int nm_version;
    unsigned int nm_flags;
    const char *nm_filename;
    napi_addon_register_func nm_register_func;
    const char *nm_modname;
    void *nm_priv;
    void *reserved[4];
} napi_module;
void napi_module_register(napi_module *mod);
#ifdef __cplusplus
}
#define NAPI_MODULE(modname, regfunc)                                                                  \
    static napi_module _module = {1, 0, __FILE__, regfunc, #modname, nullptr, {nullptr}};               \
    static void _register_##modname(void) __attribute__((constructor));                                \
    static void _register_##modname(void) { napi_module_register(&_module); }
#endif
#endif
