/*
 * Resource handle.  Replaces cpp/include/cugraph_c/resource_handle.h:19-31.
 *
 * The reference's argument is a raft::handle_t* (cpp/src/c_api/resource_handle.hpp:12-25).  raft is
 * not part of this build: NULL creates a single-GPU handle on the current device (its own stream
 * and stream-ordered memory pool); a non-NULL argument (the reference's raft handle with NCCL comms) is refused: handle
 * creation fails and returns NULL.  get_rank / get_comm_size therefore always report 0 / 1; multi-GPU runs are driven by
 * cugraph_b200/mg.py over the cugraph_b200_block_* entry points (cugraph_c/b200_ext.h).
 */
#pragma once
#include <cugraph_c/error.h>
#include <cugraph_c/export.h>
#include <cugraph_c/types.h>
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cugraph_resource_handle_ { int32_t align_; } cugraph_resource_handle_t;

CUGRAPH_EXPORT cugraph_resource_handle_t* cugraph_create_resource_handle(void* raft_handle);
CUGRAPH_EXPORT int cugraph_resource_handle_get_comm_size(const cugraph_resource_handle_t* handle);
CUGRAPH_EXPORT int cugraph_resource_handle_get_rank(const cugraph_resource_handle_t* handle);
CUGRAPH_EXPORT void cugraph_free_resource_handle(cugraph_resource_handle_t* handle);

#ifdef __cplusplus
}
#endif
