/*
 * Error codes and the opaque error object.  Replaces cpp/include/cugraph_c/error.h:15-29.
 * Convention (reference cpp/src/c_api/utils.hpp:13-47): every entry point returns a code; on
 * failure *error receives a heap object whose text is read with cugraph_error_message() and
 * released with cugraph_error_free(); on success *error is NULL.  No C++ exception crosses.
 */
#pragma once
#include <cugraph_c/export.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum cugraph_error_code_ {
  CUGRAPH_SUCCESS = 0,
  CUGRAPH_UNKNOWN_ERROR,
  CUGRAPH_INVALID_HANDLE,
  CUGRAPH_ALLOC_ERROR,
  CUGRAPH_INVALID_INPUT,
  CUGRAPH_NOT_IMPLEMENTED,
  CUGRAPH_UNSUPPORTED_TYPE_COMBINATION
} cugraph_error_code_t;

typedef struct cugraph_error_ { int32_t align_; } cugraph_error_t;

CUGRAPH_EXPORT const char* cugraph_error_message(const cugraph_error_t* error);
CUGRAPH_EXPORT void cugraph_error_free(cugraph_error_t* error);

#ifdef __cplusplus
}
#endif
