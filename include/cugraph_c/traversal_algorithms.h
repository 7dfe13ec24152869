/*
 * BFS / SSSP + the paths result.  Replaces cpp/include/cugraph_c/traversal_algorithms.h:
 * accessors :39-70, cugraph_bfs :100, cugraph_sssp :132.
 *
 * Semantics preserved (cpp/src/c_api/bfs.cpp:22-216, sssp.cpp:20-149, bfs_impl.cuh:133-869,
 * sssp_impl.cuh:169-566): sources / source are EXTERNAL ids; an id that is not a vertex gives
 * CUGRAPH_INVALID_INPUT; sources' dtype must equal the graph's vertex dtype; unreached vertices
 * get distance INT32_MAX / INT64_MAX (BFS) or FLT_MAX / DBL_MAX (SSSP) and predecessor -1;
 * direction_optimizing requires a symmetric graph; predecessors are any valid BFS / shortest-path
 * tree parent (the reference uses reduce_op::any, bfs_impl.cuh:467 — only validity is defined).
 */
#pragma once
#include <cugraph_c/error.h>
#include <cugraph_c/export.h>
#include <cugraph_c/graph.h>
#include <cugraph_c/resource_handle.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t align_; } cugraph_paths_result_t;

CUGRAPH_EXPORT cugraph_type_erased_device_array_view_t* cugraph_paths_result_get_vertices(
  cugraph_paths_result_t* result);
CUGRAPH_EXPORT cugraph_type_erased_device_array_view_t* cugraph_paths_result_get_distances(
  cugraph_paths_result_t* result);
CUGRAPH_EXPORT cugraph_type_erased_device_array_view_t* cugraph_paths_result_get_predecessors(
  cugraph_paths_result_t* result);
CUGRAPH_EXPORT void cugraph_paths_result_free(cugraph_paths_result_t* result);

CUGRAPH_EXPORT cugraph_error_code_t cugraph_bfs(
  const cugraph_resource_handle_t* handle, cugraph_graph_t* graph,
  cugraph_type_erased_device_array_view_t* sources, bool_t direction_optimizing, size_t depth_limit,
  bool_t compute_predecessors, bool_t do_expensive_check, cugraph_paths_result_t** result,
  cugraph_error_t** error);

CUGRAPH_EXPORT cugraph_error_code_t cugraph_sssp(
  const cugraph_resource_handle_t* handle, cugraph_graph_t* graph, size_t source, double cutoff,
  bool_t compute_predecessors, bool_t do_expensive_check, cugraph_paths_result_t** result,
  cugraph_error_t** error);

/* Paths of a BFS (or SSSP) result: cpp/include/cugraph_c/traversal_algorithms.h:141-201, cpp/src/c_api/extract_paths.cpp,
 * cpp/src/traversal/extract_bfs_paths_impl.cuh:129-238.  Row i of the row-major matrix holds the path
 * source ... destinations[i] in columns 0 .. distance(destinations[i]); unused cells and the rows of destinations that were
 * not reached hold the invalid vertex (-1).  max_path_length = 1 + the largest distance of a destination with a predecessor. */
typedef struct { int32_t align_; } cugraph_extract_paths_result_t;

CUGRAPH_EXPORT cugraph_error_code_t cugraph_extract_paths(
  const cugraph_resource_handle_t* handle, cugraph_graph_t* graph, const cugraph_type_erased_device_array_view_t* sources,
  const cugraph_paths_result_t* paths_result, const cugraph_type_erased_device_array_view_t* destinations,
  cugraph_extract_paths_result_t** result, cugraph_error_t** error);
CUGRAPH_EXPORT size_t cugraph_extract_paths_result_get_max_path_length(cugraph_extract_paths_result_t* result);
CUGRAPH_EXPORT cugraph_type_erased_device_array_view_t* cugraph_extract_paths_result_get_paths(
  cugraph_extract_paths_result_t* result);
CUGRAPH_EXPORT void cugraph_extract_paths_result_free(cugraph_extract_paths_result_t* result);

#ifdef __cplusplus
}
#endif
