/*
 * Graph construction / destruction.  Replaces cpp/include/cugraph_c/graph.h:23-26 (properties),
 * :69 (create_sg), :128 (create_with_times_sg — what pylibcugraph.SGGraph calls, graphs.pyx:282),
 * :177 (create_sg_from_csr), :239/:310 (create_mg / create_with_times_mg), :335 (free).
 *
 * Staging contract (reference cpp/src/c_api/graph_sg.cpp:89-330): optional self-loop removal,
 * multi-edge removal, symmetrisation; vertices renumbered to degree-descending internal ids;
 * compressed-sparse storage with per-row sorted neighbours; results of every algorithm are
 * reported against EXTERNAL ids.  This implementation keeps both orientations (CSR and CSC) on
 * demand instead of transposing in place (reference cpp/src/c_api/graph.hpp:86-150).
 * edge_ids / edge_type_ids / edge times are accepted and validated but not stored: no algorithm on
 * the PageRank/BFS/SSSP path reads them.
 */
#pragma once
#include <cugraph_c/array.h>
#include <cugraph_c/export.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t align_; } cugraph_graph_t;

typedef struct {
  bool_t is_symmetric;
  bool_t is_multigraph;
} cugraph_graph_properties_t;

CUGRAPH_EXPORT cugraph_error_code_t cugraph_graph_create_sg(
  const cugraph_resource_handle_t* handle, const cugraph_graph_properties_t* properties,
  const cugraph_type_erased_device_array_view_t* vertices,
  const cugraph_type_erased_device_array_view_t* src,
  const cugraph_type_erased_device_array_view_t* dst,
  const cugraph_type_erased_device_array_view_t* weights,
  const cugraph_type_erased_device_array_view_t* edge_ids,
  const cugraph_type_erased_device_array_view_t* edge_type_ids,
  bool_t store_transposed, bool_t renumber, bool_t drop_self_loops, bool_t drop_multi_edges,
  bool_t symmetrize, bool_t do_expensive_check, cugraph_graph_t** graph, cugraph_error_t** error);

CUGRAPH_EXPORT cugraph_error_code_t cugraph_graph_create_with_times_sg(
  const cugraph_resource_handle_t* handle, const cugraph_graph_properties_t* properties,
  const cugraph_type_erased_device_array_view_t* vertices,
  const cugraph_type_erased_device_array_view_t* src,
  const cugraph_type_erased_device_array_view_t* dst,
  const cugraph_type_erased_device_array_view_t* weights,
  const cugraph_type_erased_device_array_view_t* edge_ids,
  const cugraph_type_erased_device_array_view_t* edge_type_ids,
  const cugraph_type_erased_device_array_view_t* edge_start_time_ids,
  const cugraph_type_erased_device_array_view_t* edge_end_time_ids,
  bool_t store_transposed, bool_t renumber, bool_t drop_self_loops, bool_t drop_multi_edges,
  bool_t symmetrize, bool_t do_expensive_check, cugraph_graph_t** graph, cugraph_error_t** error);

CUGRAPH_EXPORT cugraph_error_code_t cugraph_graph_create_sg_from_csr(
  const cugraph_resource_handle_t* handle, const cugraph_graph_properties_t* properties,
  const cugraph_type_erased_device_array_view_t* offsets,
  const cugraph_type_erased_device_array_view_t* indices,
  const cugraph_type_erased_device_array_view_t* weights,
  const cugraph_type_erased_device_array_view_t* edge_ids,
  const cugraph_type_erased_device_array_view_t* edge_type_ids,
  bool_t store_transposed, bool_t renumber, bool_t symmetrize, bool_t do_expensive_check,
  cugraph_graph_t** graph, cugraph_error_t** error);

/* Multi-GPU: every rank passes its share of the edge list (num_arrays chunks, concatenated). */
CUGRAPH_EXPORT cugraph_error_code_t cugraph_graph_create_mg(
  cugraph_resource_handle_t const* handle, cugraph_graph_properties_t const* properties,
  cugraph_type_erased_device_array_view_t const* const* vertices,
  cugraph_type_erased_device_array_view_t const* const* src,
  cugraph_type_erased_device_array_view_t const* const* dst,
  cugraph_type_erased_device_array_view_t const* const* weights,
  cugraph_type_erased_device_array_view_t const* const* edge_ids,
  cugraph_type_erased_device_array_view_t const* const* edge_type_ids,
  bool_t store_transposed, size_t num_arrays, bool_t drop_self_loops, bool_t drop_multi_edges,
  bool_t symmetrize, bool_t do_expensive_check, cugraph_graph_t** graph, cugraph_error_t** error);

CUGRAPH_EXPORT cugraph_error_code_t cugraph_graph_create_with_times_mg(
  cugraph_resource_handle_t const* handle, cugraph_graph_properties_t const* properties,
  cugraph_type_erased_device_array_view_t const* const* vertices,
  cugraph_type_erased_device_array_view_t const* const* src,
  cugraph_type_erased_device_array_view_t const* const* dst,
  cugraph_type_erased_device_array_view_t const* const* weights,
  cugraph_type_erased_device_array_view_t const* const* edge_ids,
  cugraph_type_erased_device_array_view_t const* const* edge_type_ids,
  cugraph_type_erased_device_array_view_t const* const* edge_start_time_ids,
  cugraph_type_erased_device_array_view_t const* const* edge_end_time_ids,
  bool_t store_transposed, size_t num_arrays, bool_t drop_self_loops, bool_t drop_multi_edges,
  bool_t symmetrize, bool_t do_expensive_check, cugraph_graph_t** graph, cugraph_error_t** error);

CUGRAPH_EXPORT void cugraph_graph_free(cugraph_graph_t* graph);

#ifdef __cplusplus
}
#endif
