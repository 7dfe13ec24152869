/* Graph functions next to the hot path: vertex degrees.  Replaces cpp/include/cugraph_c/graph_functions.h:284-394
 * (cpp/src/c_api/degrees.cpp, degrees_result.cpp).  The other functions of that header (two-hop neighbours, induced subgraphs,
 * multi-edge counts, edge-list extraction, …) are not part of this build. */
#pragma once
#include <cugraph_c/array.h>
#include <cugraph_c/error.h>
#include <cugraph_c/export.h>
#include <cugraph_c/graph.h>
#include <cugraph_c/resource_handle.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t align_; } cugraph_degrees_result_t;

/* Degrees of `source_vertices` (NULL: of every vertex), as arrays of the graph's edge type.  in-only / out-only calls leave the
 * other accessor NULL; on a symmetric graph both accessors return the same values. */
CUGRAPH_EXPORT cugraph_error_code_t cugraph_in_degrees(
  const cugraph_resource_handle_t* handle, cugraph_graph_t* graph, const cugraph_type_erased_device_array_view_t* source_vertices,
  bool_t do_expensive_check, cugraph_degrees_result_t** result, cugraph_error_t** error);
CUGRAPH_EXPORT cugraph_error_code_t cugraph_out_degrees(
  const cugraph_resource_handle_t* handle, cugraph_graph_t* graph, const cugraph_type_erased_device_array_view_t* source_vertices,
  bool_t do_expensive_check, cugraph_degrees_result_t** result, cugraph_error_t** error);
CUGRAPH_EXPORT cugraph_error_code_t cugraph_degrees(
  const cugraph_resource_handle_t* handle, cugraph_graph_t* graph, const cugraph_type_erased_device_array_view_t* source_vertices,
  bool_t do_expensive_check, cugraph_degrees_result_t** result, cugraph_error_t** error);
CUGRAPH_EXPORT cugraph_type_erased_device_array_view_t* cugraph_degrees_result_get_vertices(cugraph_degrees_result_t* degrees_result);
CUGRAPH_EXPORT cugraph_type_erased_device_array_view_t* cugraph_degrees_result_get_in_degrees(cugraph_degrees_result_t* degrees_result);
CUGRAPH_EXPORT cugraph_type_erased_device_array_view_t* cugraph_degrees_result_get_out_degrees(cugraph_degrees_result_t* degrees_result);
CUGRAPH_EXPORT void cugraph_degrees_result_free(cugraph_degrees_result_t* degrees_result);

#ifdef __cplusplus
}
#endif
