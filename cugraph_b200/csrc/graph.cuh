// Graph data model on one B200: the graph_view_t / edge_partition_device_view_t role
// (reference cpp/include/cugraph/graph_view.hpp:840-1122, edge_partition_device_view.cuh:912-1215),
// laid out for the kernels in spmv.cu / traverse.cu.
//
// Three id spaces:
//   external : whatever the caller passed (int32 or int64)
//   rank     : dense 0..V-1 in ascending external-id order (or == external when renumber=false)
//   internal : 0..V-1 in DESCENDING degree of the primary orientation's rows, ties by rank.
// Kernels only see internal ids (always int32: V < 2^31).  The ordering *is* the degree binning:
// rows with degree >= 32 form a prefix whose edges form a prefix of `indices`
// (the role of the reference's segment offsets, graph_view.hpp:242-254 / renumber_edgelist_impl.cuh:740-828).
#pragma once
#include "common.cuh"

#include <algorithm>
#include <cstdlib>
#include <memory>

namespace b200 {

// degree thresholds that delimit the row bins (descending); bin k holds rows with
// kSegThreshold[k] <= degree < kSegThreshold[k-1].
constexpr int kNumSeg                = 7;
constexpr int kSegThreshold[kNumSeg] = {32, 16, 8, 4, 2, 1, 0};

// edges per warp / per CTA in the edge-balanced kernel for the degree>=32 prefix
constexpr int kWarpChunk  = 1024;
constexpr int kWarpsPerCta = 8;

// ---------------------------------------------------------------------------------------------
// Column-blocked copy of the degree>=32 prefix for the shared-memory gather kernel (spmv_hot.cuh).
// The source (column) space is cut into B "hot" blocks of W vertices (W * sizeof(T) = 192 KiB, the
// slice of x one CTA keeps in shared memory) plus one cold block (everything >= B*W).  Because rows
// keep their neighbours sorted by source id, a row's adjacency is already partitioned by block; the
// copy stores the segments block-major: all (row, block 0) segments, then block 1, ...
// Hot blocks store 16-bit local column ids (halves the index stream), the cold block 32-bit ids.
// Every (row, block) segment is cut into LANE SLOTS of 8 entries (16 bytes of ids: one 128-bit load per
// lane); the last slot of a segment is padded with a column that reads 0, so the kernel is predicate-free.
// ---------------------------------------------------------------------------------------------
constexpr int kHotSliceBytes = 192 * 1024;  // x slice a CTA keeps in shared memory
constexpr int kHotZeroPad    = 64;          // trailing elements of the slice that hold zeros (padding target)
constexpr int kHotSlot       = 8;           // entries per lane slot

struct hot_layout_t {
  int W{0};      // source columns per hot block (= slice elements - kHotZeroPad)
  int B{0};      // hot blocks; the cold block (sources >= B*W) is block B
  int32_t n_hi{0};    // rows covered by the layout: the prefix of degree >= 32 rows (seg_k = 0), or, experimentally,
  int64_t nnz_hi{0};  // the prefix down to a lower degree bound (seg_k > 0: rows [0, seg[seg_k]))
  int seg_k{0};
  bool bank_order{false};  // experimental: entries inside the lane slots ordered by shared-memory bank, padding on any of the zero columns
  int64_t n_hot_slots{0};
  int64_t n_slots{0};
  // A (row, block) segment is cut into PIECES of <= 64 entries = <= 8 lane slots of 8 entries.  Pieces are
  // ordered by (block, slots per piece); 32 consecutive pieces of one class form a GROUP = the work of one
  // warp (lane = piece).  Slots of a group are stored step-major: slot (step j, lane l) at group_base + 32 j + l.
  dbuf slot_idx16;   // n_hot_slots x 8 x uint16 : column - block*W, padding -> W (the slice's zero column)
  dbuf slot_idx32;   // (n_slots - n_hot_slots) x 8 x int32 : cold columns, padding -> n_vertices (x is 0 there)
  dbuf slot_w;       // n_slots x 8 x T, padding 0; or empty
  // EXPERIMENTAL narrow classes (CUGRAPH_B200_HOT_NARROW=1, unweighted graphs, hot blocks only; consumed by
  // k_spmv_blocked_x): pieces of 3-4 entries use an 8-byte slot (class code 16), of 2 entries a 4-byte slot (32), of 1 entry
  // a 2-byte slot (64)
  bool narrow{false};
  dbuf slot_idx_h;   // n_hslots x 4 x uint16
  dbuf slot_idx_q;   // n_qslots x 2 x uint16
  dbuf slot_idx_s;   // n_sslots x 1 x uint16 : one-entry pieces (class code 64)
  dbuf seg_row;      // per (group, lane): row of the piece, -1 for the unused lanes of a class's last group
  dbuf subs;         // n_subs x hot_sub_t (spmv_hot.cuh): consecutive groups of one class (cls 1..8; 16 / 32 / 64 = narrow)
  dbuf units;        // n_units x hot_unit_t: consecutive sub-units of one block, about 8192 slots
  int32_t n_subs{0};
  int32_t n_units{0};
  dbuf cta_range;    // (n_cta + 1) x int32 : CTA c owns units [cta_range[c], cta_range[c+1]) (cost-balanced)
  dbuf unit_counter; // n_cta x int : per-range cursors, also used for stealing (reset by the finish kernel)
  int n_cta{0};      // CTAs of the persistent kernel = min(SM count, n_units)
};

// EXPERIMENTAL (CUGRAPH_B200_LOW_ELL=1): exact-degree classes of the degree < 32 rows.  Rows are degree-descending, so
// the rows of degree d are the contiguous range [row_begin[d], row_begin[d] + n[d]) and their adjacency is a dense
// n[d] x d matrix starting at indices[off0[d]]; `idx` holds it TRANSPOSED (entry k of row i at off0[d] - off0[31] +
// k * n[d] + i): a lane per row reads coalesced, needs no offsets, and can own several rows.
struct low_ell_t {
  int32_t row_begin[32]{};
  int32_t n[32]{};
  long long base[32]{};  // start of class d inside idx / w
  dbuf idx;              // (nnz - nnz_hi) x int32
  dbuf w;                // same x T, or empty
};

// EXPERIMENTAL (CUGRAPH_B200_HOT_MIN_DEGREE = 32 | 16 | 8 | 4 | 2 | 1, default 32): rows down to that degree go through the
// piece layout of the blocked sweep instead of the gather kernel for low rows.  Returns the index into csx_t::seg.
inline int hot_seg_index()
{
  const char* e = std::getenv("CUGRAPH_B200_HOT_MIN_DEGREE");
  const int d   = e ? std::atoi(e) : 32;
  for (int k = 0; k < kNumSeg - 1; ++k)
    if (kSegThreshold[k] == d) return k;
  return 0;
}

// One orientation: compressed rows over `n_rows` physical rows.
// row_vertex == nullptr  -> physical row r is vertex r (rows are degree-descending by construction)
// row_vertex != nullptr  -> physical row r is vertex row_vertex[r] (a lazily built transpose whose
//                           rows were re-sorted by ITS degree so that the same kernels apply)
struct csx_t {
  int32_t n_rows{0};
  int64_t nnz{0};
  bool offs64{false};
  dbuf offsets;     // (n_rows+1) x int32|int64
  dbuf indices;     // nnz x int32, ascending within a row
  dbuf weights;     // nnz x float|double, or empty
  dbuf row_vertex;  // n_rows x int32, or empty
  bool degree_sorted{true};  // rows in descending degree (binning valid)
  int32_t seg[kNumSeg + 1]{};  // seg[k] = #rows with degree >= kSegThreshold[k]; seg[kNumSeg]=n_rows
  int64_t nnz_hi{0};           // edges in rows with degree >= 32 (= offsets[seg[0]])
  // per-warp-chunk metadata for the edge-balanced kernel over [0, nnz_hi)
  int32_t n_chunks{0};
  dbuf chunk_first_row;  // n_chunks+1 x int32 : row that contains edge c*kWarpChunk
  int32_t n_split{0};
  dbuf split_rows;  // n_split x int32 : rows that straddle a chunk boundary (each listed once)
  // lazily built column-blocked copies (float / double element width) and cached out-weight sums
  mutable std::unique_ptr<hot_layout_t> hot4, hot8;
  mutable bool hot4_tried{false}, hot8_tried{false};
  mutable dbuf out_w;  // n_vertices x T : per-source sum of edge weights (or out-degree), T = weight type
  mutable std::unique_ptr<low_ell_t> low_ell;
  mutable bool low_ell_tried{false};
};

// rows that may need an fp64 accumulator in a sweep (callers size acc_hi with it)
inline int32_t acc_rows(csx_t const& c) { return std::max(std::max(c.seg[0], c.seg[hot_seg_index()]), 1); }

struct graph_impl {
  cugraph_data_type_id_t vertex_type{INT32};
  cugraph_data_type_id_t edge_type{INT32};
  cugraph_data_type_id_t weight_type{FLOAT32};
  bool weighted{false};
  bool is_symmetric{false};
  bool is_multigraph{false};
  bool store_transposed{false};
  bool renumbered{true};
  int32_t n_vertices{0};
  int64_t n_edges{0};
  int device{0};

  // id maps
  dbuf ext_of_int;    // V x vertex_type : external id of internal vertex i   (the "number_map")
  dbuf sorted_ext;    // V x vertex_type : external ids ascending (renumber=true only)
  dbuf int_of_rank;   // V x int32       : internal id of rank r
  // renumber=false: reported order is external order; results are permuted through int_of_rank.

  std::unique_ptr<csx_t> primary;    // orientation requested at creation (degree-sorted, identity rows)
  std::unique_ptr<csx_t> pull_alt;   // lazily built CSC with re-sorted rows (PageRank on a CSR graph)
  std::unique_ptr<csx_t> push_alt;   // lazily built CSR in vertex order (BFS/SSSP on a CSC graph)

  // multi-GPU (mg.cu): this rank's blocks of the 2D partition
  void* mg{nullptr};
};

inline graph_impl* G(cugraph_graph_t* g)
{
  B200_EXPECTS(g != nullptr, CUGRAPH_INVALID_INPUT, "graph is NULL");
  return reinterpret_cast<graph_impl*>(g);
}

// Accessors that build the missing orientation on demand (graph_build.cu).
csx_t const& pull_view(handle_impl const& h, graph_impl& g);  // rows = destinations, indices = sources
// column-blocked copy for elements of `elem_size` bytes, or nullptr when the graph is too small for it
hot_layout_t const* hot_layout(handle_impl const& h, csx_t const& c, int32_t n_vertices, size_t elem_size);
// exact-degree ELL copy of the degree < 32 rows; nullptr unless CUGRAPH_B200_LOW_ELL=1 (graph_build.cu)
low_ell_t const* low_ell_layout(handle_impl const& h, csx_t const& c, size_t elem_size);
csx_t const& push_view(handle_impl const& h, graph_impl& g);  // rows = sources, vertex-indexed offsets

// external <-> internal id helpers (graph_build.cu)
// out[i] = internal id of ext[i], or -1 if ext[i] is not a vertex.
void ext_to_int(handle_impl const& h, graph_impl const& g, void const* ext, size_t n, int32_t* out);
// in-place/out-of-place: ext_out[i] = external id of internal id in[i] (in[i] < 0 stays -1)
void int_to_ext(handle_impl const& h, graph_impl const& g, int32_t const* in, size_t n, void* ext_out);
// vertices array (external ids) in reported order
dbuf reported_vertices(handle_impl const& h, graph_impl const& g);
// permute a per-vertex result from internal order into reported order (no-op copy when renumbered)
dbuf to_reported_order(handle_impl const& h, graph_impl const& g, void const* internal_vals, size_t elem_size);

// (vertex, value) pairs with external ids -> dense internal-order vector, missing = fill
template <typename T>
dbuf collect_vertex_values(handle_impl const& h, graph_impl const& g,
                           device_array_view_impl const* verts, device_array_view_impl const* vals,
                           T fill);

std::unique_ptr<csx_t> build_binned_rows(handle_impl const& h, int32_t const* major, int32_t const* minor, void const* w,
                                         cugraph_data_type_id_t wtype, int64_t n, int32_t nv);

// ---- multi-GPU hooks (mg.cu) ----
struct mg_pr_args {
  double alpha;
  double epsilon;
  size_t max_iterations;
};
void attach_comm(handle_impl* h, void* comm);
void free_mg_graph(graph_impl* g);
void mg_pagerank(handle_impl const& h, graph_impl& g, mg_pr_args const& a, centrality_result_impl& res);

}  // namespace b200
