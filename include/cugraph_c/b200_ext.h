/*
 * Extensions that have no counterpart in the reference ABI (prefixed cugraph_b200_).
 *
 *  - multi-GPU: the reference receives NCCL communicators inside a raft::handle_t built by raft-dask / MPI
 *    (python/pylibcugraph/pylibcugraph/comms/comms_wrapper.pyx:10-32, cpp/tests/utilities/mg_utilities.cpp:37-55) and keeps
 *    the 2D-partitioned blocks inside graph_t.  raft is not part of this build: cugraph_graph_create_mg and the multi-GPU
 *    algorithm entry points return CUGRAPH_NOT_IMPLEMENTED; multi-GPU PageRank is driven by the launcher
 *    (cugraph_b200/mg.py, one process per GPU over torch.distributed / NCCL) on top of the cugraph_b200_block_* device
 *    pieces declared below.
 *  - profiling hooks used by bench.py to time the dominant kernel on the handle's stream.
 */
#pragma once
#include <cugraph_c/algorithms.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Library version string and the CUDA stream of a handle (as an integer, for event timing). */
CUGRAPH_EXPORT const char* cugraph_b200_version(void);
CUGRAPH_EXPORT void* cugraph_b200_handle_stream(const cugraph_resource_handle_t* handle);

/* Number of kernels this library launched through the handle since it was created. */
CUGRAPH_EXPORT size_t cugraph_b200_handle_launch_count(const cugraph_resource_handle_t* handle);

/*
 * Benchmark hook: run `iterations` pull-SpMV sweeps (the PageRank inner kernel set, no vertex pass)
 * on the graph's pull orientation and return the average time of ONE sweep in milliseconds,
 * measured with CUDA events on the handle's stream.  x is refreshed from a fixed vector; results
 * are written to an internal buffer.  Used for roofline.achieved in bench.py.
 */
CUGRAPH_EXPORT cugraph_error_code_t cugraph_b200_time_pull_spmv(
  const cugraph_resource_handle_t* handle, cugraph_graph_t* graph, size_t iterations,
  double* ms_per_sweep, double* algorithmic_bytes_per_sweep, cugraph_error_t** error);

/*
 * ---- multi-GPU building blocks (orchestrated by cugraph_b200/mg.py over torch.distributed) ----
 * The reference keeps its 2D-partitioned edge blocks inside graph_t and runs the exchange inside the
 * prims (update_edge_src_property / per_v_transform_reduce_e MG paths).  Here the launcher owns the
 * process groups; the library provides the device pieces.  Every call below only enqueues work on
 * the handle's stream (create the handle on the caller's stream so that collectives and kernels
 * are ordered without host synchronisation).
 */
typedef struct { int32_t align_; } cugraph_b200_block_t;

CUGRAPH_EXPORT cugraph_resource_handle_t* cugraph_b200_create_resource_handle_on_stream(void* cuda_stream);
CUGRAPH_EXPORT size_t cugraph_b200_padded_elems(size_t n, size_t elem_size);
CUGRAPH_EXPORT cugraph_error_code_t cugraph_b200_block_create(
  const cugraph_resource_handle_t* handle, size_t n_rows, size_t n_cols,
  const cugraph_type_erased_device_array_view_t* rows, const cugraph_type_erased_device_array_view_t* cols,
  const cugraph_type_erased_device_array_view_t* weights, cugraph_b200_block_t** block, cugraph_error_t** error);
CUGRAPH_EXPORT void cugraph_b200_block_free(cugraph_b200_block_t* block);
CUGRAPH_EXPORT size_t cugraph_b200_block_span(const cugraph_b200_block_t* block);
/* y[row] = alpha * sum over the block's edges (row, col) of x[col] * w; rows without edges get 0.  The FIRST sweep of a block
 * into a given y array writes every row slot; later sweeps into the same array only rewrite the rows that have edges (in a 2D
 * block most row slots are empty) — the caller must leave the other entries alone, or pass a zero-initialised array. */
CUGRAPH_EXPORT cugraph_error_code_t cugraph_b200_block_pull_sweep(
  const cugraph_resource_handle_t* handle, cugraph_b200_block_t* block,
  const cugraph_type_erased_device_array_view_t* x, cugraph_type_erased_device_array_view_t* y, double alpha,
  cugraph_error_t** error);
CUGRAPH_EXPORT cugraph_error_code_t cugraph_b200_pagerank_vertex_step(
  const cugraph_resource_handle_t* handle, const cugraph_type_erased_device_array_view_t* y,
  cugraph_type_erased_device_array_view_t* pr, const cugraph_type_erased_device_array_view_t* out_w,
  cugraph_type_erased_device_array_view_t* x, size_t n_local, double alpha, double n_vertices_global, bool_t first,
  const double* totals_prev_device, double* partial_out_device, cugraph_error_t** error);

/* RMAT edge list written into caller-allocated INT32 arrays (the role of cugraph_generate_rmat_edgelist,
 * cpp/include/cugraph_c/graph_generators.h, without its rng_state / coo objects): the reference's sampling rule, clip-and-flip
 * and id scramble (generate_rmat_edgelist.cuh:66-108, scramble.cuh:44-67) over a counter-based uniform stream, so that
 * the same (scale, seed) gives the same edges on any grid size — restated in numpy by oracle/rmat.py for the parity test. */
CUGRAPH_EXPORT cugraph_error_code_t cugraph_b200_generate_rmat_edgelist(
  const cugraph_resource_handle_t* handle, size_t scale, size_t num_edges, double a, double b, double c, uint64_t seed,
  bool_t clip_and_flip, bool_t scramble_vertex_ids, cugraph_type_erased_device_array_view_t* src,
  cugraph_type_erased_device_array_view_t* dst, cugraph_error_t** error);

/* Uniform values for synthetic edge weights / types (what cugraph_generate_edge_weights / _edge_types draw from raft's RNG):
 * out[i] = lo + u_i * (hi - lo), u_i from a 64-bit mix of (seed, i); INT32 arrays get integers in [lo, hi). */
CUGRAPH_EXPORT cugraph_error_code_t cugraph_b200_generate_uniform(const cugraph_resource_handle_t* handle, uint64_t seed, double lo,
                                                                  double hi, cugraph_type_erased_device_array_view_t* out,
                                                                  cugraph_error_t** error);

/* One level of multi-GPU BFS on this GPU's edge block (pull direction; the role of the bottom-up step of
 * cpp/src/traversal/bfs_impl.cuh:593-869 on one edge partition).  frontier_cols / visited_rows: byte flags over the block's
 * column (source) / row (destination) slots, gathered by the launcher inside the column / row group.  cand (INT64, one per row
 * slot) receives, for every unvisited row with a source in the frontier, that source's global code
 * ((col / maxpart) * grid_cols + grid_c) * maxpart + col % maxpart  (= owner rank * maxpart + local id), else -1. */
CUGRAPH_EXPORT cugraph_error_code_t cugraph_b200_block_bfs_pull(
  const cugraph_resource_handle_t* handle, cugraph_b200_block_t* block,
  const cugraph_type_erased_device_array_view_t* frontier_cols, const cugraph_type_erased_device_array_view_t* visited_rows,
  size_t maxpart, int grid_cols, int grid_c, cugraph_type_erased_device_array_view_t* cand, cugraph_error_t** error);

/* Debug hook: one sweep as PageRank would run it on this graph (the shared-memory piece stream when the graph has one)
 * against the plain sweep (an independent implementation) on the same pseudo-random x.  out[0..3] = degree >= 32 rows
 * {max relative difference, its row, that row's degree, rows above 1e-5}; out[4..7] = the same for the degree < 32 rows. */
CUGRAPH_EXPORT cugraph_error_code_t cugraph_b200_debug_compare_sweeps(const cugraph_resource_handle_t* handle,
                                                                     cugraph_graph_t* graph, double* out,
                                                                     cugraph_error_t** error);

/* Host-only planner of the sweep's work structure (chunks, phases, per-CTA phase ranges) from the piece counts per
 * (block, kind); the function graph staging itself uses.  Needs no GPU: exposed so that the host logic is testable on CPU
 * (tests/test_sweep_plan_cpu.py).  class_start has n_blocks * 11 + 1 entries (kinds S, Q, H, F1..F8).
 * Outputs: totals[3] = {step-rows, row slots, CTAs}; chunks / fills / phases are 4 x int32 records
 * ({sr_begin,row_begin,n_groups,kind}, {piece_begin,piece_end,block,0}, {block,chunk_begin,chunk_end,0});
 * cta_phase has totals[2] + 1 entries.  Returns CUGRAPH_INVALID_INPUT when a capacity is too small. */
CUGRAPH_EXPORT cugraph_error_code_t cugraph_b200_debug_plan_sweep(const int32_t* class_start, int n_blocks, int sm_count,
                                                                 int64_t* totals, int32_t* chunks, int32_t* fills,
                                                                 size_t chunks_capacity, size_t* n_chunks, int32_t* phases,
                                                                 size_t phases_capacity, size_t* n_phases, int32_t* cta_phase,
                                                                 size_t cta_capacity, cugraph_error_t** error);

#ifdef __cplusplus
}
#endif
