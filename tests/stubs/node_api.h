// Minimal DECLARATION-ONLY stand-in for Node's node_api.h (absent from this image): lets tests/test_napi_shim.py type-check
// fspt_b200/napi/fspt_napi.cc against include/fspt_b200.h.  Nothing links against it.
#pragma once
#include <stddef.h>
#include <stdint.h>
typedef struct napi_env__* napi_env; typedef struct napi_value__* napi_value; typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok } napi_status;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array } napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env, napi_callback_info);
typedef void (*napi_finalize)(napi_env, void*, void*);
#define NAPI_AUTO_LENGTH SIZE_MAX
extern "C" {
napi_status napi_throw_error(napi_env, const char*, const char*);
napi_status napi_get_value_external(napi_env, napi_value, void**);
napi_status napi_has_named_property(napi_env, napi_value, const char*, bool*);
napi_status napi_get_named_property(napi_env, napi_value, const char*, napi_value*);
napi_status napi_set_named_property(napi_env, napi_value, const char*, napi_value);
napi_status napi_get_typedarray_info(napi_env, napi_value, napi_typedarray_type*, size_t*, void**, napi_value*, size_t*);
napi_status napi_get_value_double(napi_env, napi_value, double*);
napi_status napi_get_value_int32(napi_env, napi_value, int32_t*);
napi_status napi_get_value_uint32(napi_env, napi_value, uint32_t*);
napi_status napi_get_element(napi_env, napi_value, uint32_t, napi_value*);
napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t*, napi_value*, napi_value*, void**);
napi_status napi_create_external(napi_env, void*, napi_finalize, void*, napi_value*);
napi_status napi_create_arraybuffer(napi_env, size_t, void**, napi_value*);
napi_status napi_create_typedarray(napi_env, napi_typedarray_type, size_t, napi_value, size_t, napi_value*);
napi_status napi_create_object(napi_env, napi_value*);
napi_status napi_create_int32(napi_env, int32_t, napi_value*);
napi_status napi_create_function(napi_env, const char*, size_t, napi_callback, void*, napi_value*);
}
#define NAPI_MODULE(name, init) napi_value fspt_napi_register(napi_env e, napi_value x) { return init(e, x); }
#define NODE_GYP_MODULE_NAME fspt_napi
