/*
 * node_api_min.h -- the subset of Node-API (N-API, ABI-stable since Node 8/10) that fa_napi.c uses.
 *
 * This image has no Node.js and no node_api.h, so the addon is compile-checked against these declarations
 * (they follow the documented, ABI-stable C signatures of https://nodejs.org/api/n-api.html).  The symbols are
 * resolved by the node executable when the addon is loaded; when building where Node exists, node-gyp's own
 * <node_api.h> is used instead (binding.gyp defines FA_USE_SYSTEM_NODE_API).
 */
#ifndef FA_NODE_API_MIN_H_
#define FA_NODE_API_MIN_H_
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_deferred__* napi_deferred;
typedef struct napi_async_work__* napi_async_work;
typedef struct napi_callback_info__* napi_callback_info;
typedef struct napi_ref__* napi_ref;

typedef enum { napi_ok = 0, napi_invalid_arg, napi_object_expected, napi_string_expected, napi_name_expected,
               napi_function_expected, napi_number_expected, napi_boolean_expected, napi_array_expected,
               napi_generic_failure, napi_pending_exception, napi_cancelled } napi_status;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object,
               napi_function, napi_external, napi_bigint } napi_valuetype;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array,
               napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array,
               napi_biguint64_array } napi_typedarray_type;

typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
typedef void (*napi_async_execute_callback)(napi_env env, void* data);
typedef void (*napi_async_complete_callback)(napi_env env, napi_status status, void* data);

#define NAPI_AUTO_LENGTH SIZE_MAX
#define NAPI_EXTERN
#ifdef __cplusplus
extern "C" {
#endif
napi_status napi_create_function(napi_env, const char* utf8name, size_t length, napi_callback cb, void* data, napi_value* result);
napi_status napi_set_named_property(napi_env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_get_named_property(napi_env, napi_value object, const char* utf8name, napi_value* result);
napi_status napi_has_named_property(napi_env, napi_value object, const char* utf8name, bool* result);
napi_status napi_get_cb_info(napi_env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_typeof(napi_env, napi_value value, napi_valuetype* result);
napi_status napi_get_value_double(napi_env, napi_value value, double* result);
napi_status napi_get_value_int32(napi_env, napi_value value, int32_t* result);
napi_status napi_get_value_bool(napi_env, napi_value value, bool* result);
napi_status napi_create_double(napi_env, double value, napi_value* result);
napi_status napi_create_int32(napi_env, int32_t value, napi_value* result);
napi_status napi_create_string_utf8(napi_env, const char* str, size_t length, napi_value* result);
napi_status napi_create_object(napi_env, napi_value* result);
napi_status napi_create_array_with_length(napi_env, size_t length, napi_value* result);
napi_status napi_set_element(napi_env, napi_value object, uint32_t index, napi_value value);
napi_status napi_is_typedarray(napi_env, napi_value value, bool* result);
napi_status napi_get_typedarray_info(napi_env, napi_value typedarray, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_create_arraybuffer(napi_env, size_t byte_length, void** data, napi_value* result);
napi_status napi_create_typedarray(napi_env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset,
                                   napi_value* result);
napi_status napi_create_external(napi_env, void* data, napi_finalize finalize_cb, void* finalize_hint, napi_value* result);
napi_status napi_get_value_external(napi_env, napi_value value, void** result);
napi_status napi_throw_error(napi_env, const char* code, const char* msg);
napi_status napi_create_promise(napi_env, napi_deferred* deferred, napi_value* promise);
napi_status napi_resolve_deferred(napi_env, napi_deferred deferred, napi_value resolution);
napi_status napi_reject_deferred(napi_env, napi_deferred deferred, napi_value rejection);
napi_status napi_create_async_work(napi_env, napi_value async_resource, napi_value async_resource_name,
                                   napi_async_execute_callback execute, napi_async_complete_callback complete, void* data,
                                   napi_async_work* result);
napi_status napi_queue_async_work(napi_env, napi_async_work work);
napi_status napi_delete_async_work(napi_env, napi_async_work work);
napi_status napi_create_reference(napi_env, napi_value value, uint32_t initial_refcount, napi_ref* result);
napi_status napi_delete_reference(napi_env, napi_ref ref);
#ifdef __cplusplus
}
#endif
#endif
