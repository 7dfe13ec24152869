/*
 * Type-erased device / host arrays and non-owning views.
 * Replaces cpp/include/cugraph_c/array.h:43-326 (implementation contract: cpp/src/c_api/array.hpp:17-97).
 * Views are {pointer, element count, dtype}; they never own memory.  `*_array_t` own their buffer.
 */
#pragma once
#include <cugraph_c/export.h>
#include <cugraph_c/resource_handle.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t align_; } cugraph_type_erased_device_array_t;
typedef struct { int32_t align_; } cugraph_type_erased_device_array_view_t;
typedef struct { int32_t align_; } cugraph_type_erased_host_array_t;
typedef struct { int32_t align_; } cugraph_type_erased_host_array_view_t;

/* ---- device arrays (array.h:43-121) ---- */
CUGRAPH_EXPORT cugraph_error_code_t cugraph_type_erased_device_array_create(
  const cugraph_resource_handle_t* handle, size_t n_elems, cugraph_data_type_id_t dtype,
  cugraph_type_erased_device_array_t** array, cugraph_error_t** error);
CUGRAPH_EXPORT cugraph_error_code_t cugraph_type_erased_device_array_create_from_view(
  const cugraph_resource_handle_t* handle, const cugraph_type_erased_device_array_view_t* view,
  cugraph_type_erased_device_array_t** array, cugraph_error_t** error);
CUGRAPH_EXPORT void cugraph_type_erased_device_array_free(cugraph_type_erased_device_array_t* p);
CUGRAPH_EXPORT cugraph_type_erased_device_array_view_t* cugraph_type_erased_device_array_view(
  cugraph_type_erased_device_array_t* array);
CUGRAPH_EXPORT cugraph_error_code_t cugraph_type_erased_device_array_view_as_type(
  cugraph_type_erased_device_array_t* array, cugraph_data_type_id_t dtype,
  cugraph_type_erased_device_array_view_t** result_view, cugraph_error_t** error);

/* ---- device views (array.h:123-170) ---- */
CUGRAPH_EXPORT cugraph_type_erased_device_array_view_t* cugraph_type_erased_device_array_view_create(
  void* pointer, size_t n_elems, cugraph_data_type_id_t dtype);
CUGRAPH_EXPORT void cugraph_type_erased_device_array_view_free(cugraph_type_erased_device_array_view_t* p);
CUGRAPH_EXPORT size_t cugraph_type_erased_device_array_view_size(const cugraph_type_erased_device_array_view_t* p);
CUGRAPH_EXPORT cugraph_data_type_id_t cugraph_type_erased_device_array_view_type(
  const cugraph_type_erased_device_array_view_t* p);
CUGRAPH_EXPORT const void* cugraph_type_erased_device_array_view_pointer(
  const cugraph_type_erased_device_array_view_t* p);

/* ---- host arrays and views (array.h:172-262) ---- */
CUGRAPH_EXPORT cugraph_error_code_t cugraph_type_erased_host_array_create(
  const cugraph_resource_handle_t* handle, size_t n_elems, cugraph_data_type_id_t dtype,
  cugraph_type_erased_host_array_t** array, cugraph_error_t** error);
CUGRAPH_EXPORT void cugraph_type_erased_host_array_free(cugraph_type_erased_host_array_t* p);
CUGRAPH_EXPORT cugraph_type_erased_host_array_view_t* cugraph_type_erased_host_array_view(
  cugraph_type_erased_host_array_t* array);
CUGRAPH_EXPORT cugraph_type_erased_host_array_view_t* cugraph_type_erased_host_array_view_create(
  void* pointer, size_t n_elems, cugraph_data_type_id_t dtype);
CUGRAPH_EXPORT void cugraph_type_erased_host_array_view_free(cugraph_type_erased_host_array_view_t* p);
CUGRAPH_EXPORT size_t cugraph_type_erased_host_array_size(const cugraph_type_erased_host_array_view_t* p);
CUGRAPH_EXPORT cugraph_data_type_id_t cugraph_type_erased_host_array_type(
  const cugraph_type_erased_host_array_view_t* p);
CUGRAPH_EXPORT void* cugraph_type_erased_host_array_pointer(const cugraph_type_erased_host_array_view_t* p);

/* ---- copies (array.h:264-326); all synchronous w.r.t. the handle's stream on return ---- */
CUGRAPH_EXPORT cugraph_error_code_t cugraph_type_erased_host_array_view_copy(
  const cugraph_resource_handle_t* handle, cugraph_type_erased_host_array_view_t* dst,
  const cugraph_type_erased_host_array_view_t* src, cugraph_error_t** error);
CUGRAPH_EXPORT cugraph_error_code_t cugraph_type_erased_device_array_view_copy_from_host(
  const cugraph_resource_handle_t* handle, cugraph_type_erased_device_array_view_t* dst,
  const byte_t* h_src, cugraph_error_t** error);
CUGRAPH_EXPORT cugraph_error_code_t cugraph_type_erased_device_array_view_copy_to_host(
  const cugraph_resource_handle_t* handle, byte_t* h_dst,
  const cugraph_type_erased_device_array_view_t* src, cugraph_error_t** error);
CUGRAPH_EXPORT cugraph_error_code_t cugraph_type_erased_device_array_view_copy(
  const cugraph_resource_handle_t* handle, cugraph_type_erased_device_array_view_t* dst,
  const cugraph_type_erased_device_array_view_t* src, cugraph_error_t** error);

#ifdef __cplusplus
}
#endif
