/*
 * Error reporting of the C ABI.  Replaces cpp/include/cugraph_c/error.h:15-29 (codes and the two accessors are the
 * reference's ABI; the notes are about this implementation).
 *
 * Contract (reference cpp/src/c_api/utils.hpp:13-47): every entry point returns a code and takes a cugraph_error_t**.
 * On success *error is NULL.  On failure *error is a heap object owned by the caller: read its text with
 * cugraph_error_message(), release it with cugraph_error_free().  No C++ exception ever crosses the boundary; inside the
 * library failures are b200::capi_exception (csrc/common.cuh) and the `guarded` wrapper translates them.
 */
#pragma once
#include <cugraph_c/export.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum cugraph_error_code_ {
  CUGRAPH_SUCCESS                      = 0,
  CUGRAPH_UNKNOWN_ERROR                = 1, /* argument checks of the algorithms themselves (alpha range, non-convergence,
                                               CUDA runtime failures): the reference reports those as exceptions' what() */
  CUGRAPH_INVALID_HANDLE               = 2, /* NULL resource handle */
  CUGRAPH_ALLOC_ERROR                  = 3, /* cudaErrorMemoryAllocation / std::bad_alloc */
  CUGRAPH_INVALID_INPUT                = 4, /* NULL out-pointers, mismatched array types or sizes, ids that are no vertices */
  CUGRAPH_NOT_IMPLEMENTED              = 5, /* the raft-comms multi-GPU constructors; see b200_ext.h for the multi-GPU path */
  CUGRAPH_UNSUPPORTED_TYPE_COMBINATION = 6  /* e.g. INT32 vertices with INT64 edge ids */
} cugraph_error_code_t;

/* opaque: only ever handled through a pointer */
typedef struct cugraph_error_ {
  int32_t align_;
} cugraph_error_t;

/* text of a failure; valid until cugraph_error_free(error).  NULL-safe: returns NULL for NULL. */
CUGRAPH_EXPORT const char* cugraph_error_message(const cugraph_error_t* error);

/* releases the object; NULL is a no-op */
CUGRAPH_EXPORT void cugraph_error_free(cugraph_error_t* error);

#ifdef __cplusplus
}
#endif
