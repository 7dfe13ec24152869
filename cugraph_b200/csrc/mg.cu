// Multi-GPU building blocks (one process per GPU).  The 2D edge partition, the vertex -> GPU map and
// the collectives (all-gather of x over the column group, reduce-scatter of partial y over the row
// group — the roles of update_edge_src_property's grouped ncclBroadcast and
// per_v_transform_reduce_e's grouped ncclReduce, update_edge_src_dst_property.cuh:550-579 /
// per_v_transform_reduce_e.cuh:3389-3407) are orchestrated by cugraph_b200/mg.py over
// torch.distributed (NCCL on NVLink 5 / NVSwitch).  This file provides the device-side pieces behind
// the C ABI: a resource handle bound to the caller's CUDA stream, rectangular edge blocks with the
// same binned / column-blocked layout as the single-GPU graph, the block pull sweep and the fused
// per-iteration vertex step.  All calls only ENQUEUE work on the handle's stream.
#include "sweep.cuh"

#include <algorithm>

namespace b200 {

struct block_impl {
  std::unique_ptr<csx_t> csx;
  int32_t n_rows{0}, n_cols{0}, n_span{0};
  cugraph_data_type_id_t wtype{FLOAT32};
  dbuf acc_hi;
  dbuf state;  // pr_state_t with init = 0, done = 0
  // the y array whose rows WITHOUT edges this block has already written (0): later sweeps into the same array only finish the
  // rows that have edges — in a 2D block more than half of the row slots are empty (the caller must not write them either)
  void const* y_complete{nullptr};
};

namespace {

__device__ __forceinline__ double block_sum2(double v, double* smem)
{
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < (blockDim.x >> 5)) ? smem[threadIdx.x] : 0.0;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;
}

// owner slice, one PageRank iteration (pagerank_impl.cuh:225-251, 311-318 fused):
//   init    = (dangling_prev * alpha + 1 - alpha) / V        (dangling_prev from totals_prev[1])
//   pr_new  = first ? pr : y + init ; diff += |pr_new - pr| ; dangling += pr_new where out_w == 0
//   x       = pr_new / (out_w or 1) ; pr = pr_new
template <typename T>
__global__ void __launch_bounds__(256)
k_mg_vertex_step(T const* __restrict__ y, T* __restrict__ pr, T const* __restrict__ out_w, T* __restrict__ x, int32_t n,
                 double alpha, double n_vertices_global, int first, double const* __restrict__ totals_prev,
                 double* __restrict__ partial_out)
{
  __shared__ double smem[8];
  const double init = first ? 0.0 : (totals_prev[1] * alpha + (1.0 - alpha)) / n_vertices_global;
  double diff = 0.0, dang = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const T old = pr[i];
    const T nv  = first ? old : (T)((double)y[i] + init);
    const T ow  = out_w[i];
    diff += fabs((double)nv - (double)old);
    if (ow == (T)0) dang += (double)nv;
    x[i]  = (ow == (T)0) ? nv : nv / ow;
    pr[i] = nv;
  }
  diff = block_sum2(diff, smem);
  dang = block_sum2(dang, smem);
  if (threadIdx.x == 0) {
    atomicAdd(partial_out + 0, diff);
    atomicAdd(partial_out + 1, dang);
  }
}


// ---- one level of multi-GPU BFS on this GPU's edge block, pull direction (the MG form of k_bfs_bottomup, traverse.cu;
// reference: the bottom-up step of bfs_impl.cuh:593-869 on an edge partition, with the frontier arriving through
// fill_edge_dst_property-style broadcasts, fill_edge_src_dst_property.cuh:1368).  The block stores its edges by destination
// slot (rows) with the source slots as neighbours (columns, ascending).  frontier[col] / visited[row] are byte flags over
// the block's column / row slots (the launcher all-gathers them inside the column / row group); every unvisited row scans
// its sources until it meets one in the frontier and reports it as cand[row] = GLOBAL code of that source
// ((owner rank) * maxpart + local id, owner rank = (col / maxpart) * grid_cols + grid_c), else -1.  Rows of degree >= 32
// (the prefix of the degree-ordered physical rows) take a warp each with a ballot early exit, the others a thread each.
template <typename O>
__global__ void __launch_bounds__(256)
k_block_bfs_pull_hi(O const* __restrict__ off, int32_t const* __restrict__ idx, int32_t const* __restrict__ row_vertex, int32_t n_hi,
                    uint8_t const* __restrict__ frontier, uint8_t const* __restrict__ visited, long long maxpart, int grid_cols,
                    int grid_c, long long* __restrict__ cand)
{
  const int lane = threadIdx.x & 31;
  for (long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; r < n_hi; r += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int slot = row_vertex ? row_vertex[r] : (int)r;
    if (visited[slot]) continue;
    const long long e1 = (long long)off[r + 1];
    int found          = -1;
    for (long long e = (long long)off[r] + lane; __any_sync(0xffffffffu, e < e1); e += 32) {
      const int col = e < e1 ? idx[e] : -1;
      const bool hit = col >= 0 && frontier[col] != 0;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) {
        found = __shfl_sync(0xffffffffu, col, __ffs((int)m) - 1);
        break;
      }
    }
    if (lane == 0 && found >= 0) cand[slot] = ((long long)(found / maxpart) * grid_cols + grid_c) * maxpart + (found % maxpart);
  }
}

template <typename O>
__global__ void __launch_bounds__(256)
k_block_bfs_pull_low(O const* __restrict__ off, int32_t const* __restrict__ idx, int32_t const* __restrict__ row_vertex, int32_t r0,
                     int32_t r1, uint8_t const* __restrict__ frontier, uint8_t const* __restrict__ visited, long long maxpart,
                     int grid_cols, int grid_c, long long* __restrict__ cand)
{
  for (long long r = r0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; r < r1; r += (long long)gridDim.x * blockDim.x) {
    const int slot = row_vertex ? row_vertex[r] : (int)r;
    if (visited[slot]) continue;
    const long long e1 = (long long)off[r + 1];
    for (long long e = (long long)off[r]; e < e1; ++e) {
      const int col = idx[e];
      if (frontier[col]) {
        cand[slot] = ((long long)(col / maxpart) * grid_cols + grid_c) * maxpart + (col % maxpart);
        break;
      }
    }
  }
}

__global__ void k_fill_i64(long long* __restrict__ a, long long n, long long v)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] = v;
}

template <typename O>
void block_bfs_pull(handle_impl const& h, csx_t const& c, uint8_t const* frontier, uint8_t const* visited, long long maxpart,
                    int grid_cols, int grid_c, long long* cand, int32_t n_row_slots)
{
  B200_LAUNCH(h, k_fill_i64, std::min((n_row_slots + 255) / 256 + 1, h.sm_count * 8), 256, 0, cand, (long long)n_row_slots, -1ll);
  const int32_t n_hi = c.degree_sorted ? c.seg[0] : 0;
  const int32_t n_ne = c.degree_sorted ? c.seg[kNumSeg - 2] : c.n_rows;  // rows with at least one edge
  if (n_hi > 0)
    B200_LAUNCH(h, (k_block_bfs_pull_hi<O>), std::min((n_hi + 7) / 8, h.sm_count * 16), 256, 0, c.offsets.as<O>(),
                c.indices.as<int32_t>(), c.row_vertex.as<int32_t>(), n_hi, frontier, visited, maxpart, grid_cols, grid_c, cand);
  if (n_ne > n_hi)
    B200_LAUNCH(h, (k_block_bfs_pull_low<O>), std::min((n_ne - n_hi + 255) / 256, h.sm_count * 16), 256, 0, c.offsets.as<O>(),
                c.indices.as<int32_t>(), c.row_vertex.as<int32_t>(), n_hi, n_ne, frontier, visited, maxpart, grid_cols, grid_c, cand);
}

}  // namespace

void attach_comm(handle_impl*, void*)
{
  throw capi_exception(CUGRAPH_NOT_IMPLEMENTED,
                       "multi-GPU goes through cugraph_b200.mg (torch.distributed) + the cugraph_b200_block_* entry points");
}

void free_mg_graph(graph_impl*) {}

void mg_pagerank(handle_impl const&, graph_impl&, mg_pr_args const&, centrality_result_impl&)
{
  throw capi_exception(CUGRAPH_NOT_IMPLEMENTED, "multi-GPU PageRank: use cugraph_b200.mg.MGGraph / mg.pagerank");
}

}  // namespace b200

using namespace b200;

extern "C" {

cugraph_resource_handle_t* cugraph_b200_create_resource_handle_on_stream(void* cuda_stream)
{
  try {
    auto* h = new handle_impl{};
    h->tune = tuning_t::from_env();
    CUDA_TRY(cudaGetDevice(&h->device));
    h->stream         = reinterpret_cast<cudaStream_t>(cuda_stream);
    h->borrowed_stream = true;
    CUDA_TRY(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
    register_stream(h->stream);
    register_stream(h->aux_stream);
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_a, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_b, cudaEventDisableTiming));
    cudaDeviceProp prop{};
    CUDA_TRY(cudaGetDeviceProperties(&prop, h->device));
    h->sm_count = prop.multiProcessorCount;
    h->l2_bytes = static_cast<size_t>(prop.l2CacheSize);
    CUDA_TRY(cudaMallocHost(&h->pinned, 4096));
    cudaMemPool_t pool;
    CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, h->device));
    uint64_t threshold = UINT64_MAX;
    CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    return reinterpret_cast<cugraph_resource_handle_t*>(h);
  } catch (std::exception const& e) {
    std::fprintf(stderr, "cugraph_b200_create_resource_handle_on_stream: %s\n", e.what());
    return nullptr;
  }
}

size_t cugraph_b200_padded_elems(size_t n, size_t elem_size) { return padded_x_elems((int32_t)n, elem_size); }

cugraph_error_code_t cugraph_b200_block_create(const cugraph_resource_handle_t* handle, size_t n_rows, size_t n_cols,
                                               const cugraph_type_erased_device_array_view_t* rows,
                                               const cugraph_type_erased_device_array_view_t* cols,
                                               const cugraph_type_erased_device_array_view_t* weights,
                                               cugraph_b200_block_t** block, cugraph_error_t** error)
{
  return guarded(error, [&] {
    auto const& h = H(handle);
    B200_EXPECTS(block && rows && cols, CUGRAPH_INVALID_INPUT, "NULL argument");
    auto const* r = V(rows);
    auto const* c = V(cols);
    auto const* w = V(weights);
    B200_EXPECTS(r->type == INT32 && c->type == INT32 && r->size == c->size, CUGRAPH_INVALID_INPUT,
                 "block rows / cols must be INT32 arrays of equal size");
    B200_EXPECTS(w == nullptr || ((w->type == FLOAT32 || w->type == FLOAT64) && w->size == r->size), CUGRAPH_INVALID_INPUT,
                 "block weights must be FLOAT32 / FLOAT64 with one value per edge");
    B200_EXPECTS(n_rows < (1u << 31) && n_cols < (1u << 31), CUGRAPH_INVALID_INPUT, "block too large");
    auto b     = std::make_unique<block_impl>();
    b->n_rows  = (int32_t)n_rows;
    b->n_cols  = (int32_t)n_cols;
    b->n_span  = (int32_t)std::max(n_rows, n_cols);
    b->wtype   = w ? w->type : FLOAT32;
    b->csx     = build_binned_rows(h, (int32_t const*)r->data, (int32_t const*)c->data, w ? w->data : nullptr, b->wtype,
                                   (int64_t)r->size, b->n_span);
    b->acc_hi  = make_dbuf<double>(acc_rows(*b->csx), h.stream);
    CUDA_TRY(cudaMemsetAsync(b->acc_hi.data(), 0, sizeof(double) * acc_rows(*b->csx), h.stream));
    b->state = make_dbuf<pr_state_t>(1, h.stream);
    CUDA_TRY(cudaMemsetAsync(b->state.data(), 0, sizeof(pr_state_t), h.stream));
    // build the column-blocked copy now (it is lazily created otherwise, inside the first timed sweep)
    if (!b->csx->offs64) (void)sweep_layout(h, *b->csx, b->n_span, b->wtype == FLOAT64 ? 8 : 4);
    sync(h);
    *block = reinterpret_cast<cugraph_b200_block_t*>(b.release());
  });
}

void cugraph_b200_block_free(cugraph_b200_block_t* block)
{
  if (block) delete reinterpret_cast<block_impl*>(block);
}

size_t cugraph_b200_block_span(const cugraph_b200_block_t* block)
{
  return block ? (size_t) reinterpret_cast<block_impl const*>(block)->n_span : 0;
}

// y[row] = alpha * sum_{edges (row, col)} x[col] * w ; rows without edges get 0.  Asynchronous.
cugraph_error_code_t cugraph_b200_block_pull_sweep(const cugraph_resource_handle_t* handle, cugraph_b200_block_t* block,
                                                   const cugraph_type_erased_device_array_view_t* x,
                                                   cugraph_type_erased_device_array_view_t* y, double alpha,
                                                   cugraph_error_t** error)
{
  return guarded(error, [&] {
    auto const& h = H(handle);
    B200_EXPECTS(block && x && y, CUGRAPH_INVALID_INPUT, "NULL argument");
    auto* b        = reinterpret_cast<block_impl*>(block);
    auto const* xv = V(x);
    auto const* yv = V(y);
    const bool f32 = b->wtype == FLOAT32;
    B200_EXPECTS(xv->type == b->wtype && yv->type == b->wtype, CUGRAPH_INVALID_INPUT, "x / y dtype must match the block");
    B200_EXPECTS(xv->size >= padded_x_elems(b->n_span, f32 ? 4 : 8), CUGRAPH_INVALID_INPUT,
                 "x must hold cugraph_b200_padded_elems(span) elements");
    B200_EXPECTS(yv->size >= (size_t)b->n_span, CUGRAPH_INVALID_INPUT, "y must hold `span` elements");
    csx_t const& c = *b->csx;
    auto* st       = b->state.as<pr_state_t>();
    const bool covered_only = b->y_complete == yv->data;  // the empty rows of this y hold their zeros from an earlier sweep
    if (f32) {
      if (c.offs64) launch_pull_sweep<int64_t, float>(h, c, (float const*)xv->data, (float*)yv->data, b->acc_hi.as<double>(), alpha, st);
      else launch_pull_sweep_auto<int32_t, float>(h, c, b->n_span, (float const*)xv->data, (float*)yv->data, b->acc_hi.as<double>(), alpha, st, true, covered_only);
    } else {
      if (c.offs64) launch_pull_sweep<int64_t, double>(h, c, (double const*)xv->data, (double*)yv->data, b->acc_hi.as<double>(), alpha, st);
      else launch_pull_sweep_auto<int32_t, double>(h, c, b->n_span, (double const*)xv->data, (double*)yv->data, b->acc_hi.as<double>(), alpha, st, true, covered_only);
    }
    b->y_complete = yv->data;
    check_last("block_pull_sweep");
  });
}

cugraph_error_code_t cugraph_b200_pagerank_vertex_step(const cugraph_resource_handle_t* handle,
                                                       const cugraph_type_erased_device_array_view_t* y,
                                                       cugraph_type_erased_device_array_view_t* pr,
                                                       const cugraph_type_erased_device_array_view_t* out_w,
                                                       cugraph_type_erased_device_array_view_t* x, size_t n_local,
                                                       double alpha, double n_vertices_global, bool_t first,
                                                       const double* totals_prev_device, double* partial_out_device,
                                                       cugraph_error_t** error)
{
  return guarded(error, [&] {
    auto const& h = H(handle);
    B200_EXPECTS(y && pr && out_w && x && partial_out_device, CUGRAPH_INVALID_INPUT, "NULL argument");
    auto const* yv = V(y);
    auto const* pv = V(pr);
    auto const* ov = V(out_w);
    auto const* xv = V(x);
    B200_EXPECTS(pv->type == yv->type && ov->type == yv->type && xv->type == yv->type, CUGRAPH_INVALID_INPUT, "dtype mismatch");
    B200_EXPECTS(yv->size >= n_local && pv->size >= n_local && ov->size >= n_local && xv->size >= n_local,
                 CUGRAPH_INVALID_INPUT, "arrays shorter than n_local");
    if (n_local == 0) return;
    const int grid = (int)std::min<size_t>((n_local + 255) / 256, (size_t)h.sm_count * 8);
    if (yv->type == FLOAT32)
      B200_LAUNCH(h, (k_mg_vertex_step<float>), grid, 256, 0, (float const*)yv->data, (float*)pv->data, (float const*)ov->data,
                  (float*)xv->data, (int32_t)n_local, alpha, n_vertices_global, first == TRUE ? 1 : 0, totals_prev_device,
                  partial_out_device);
    else
      B200_LAUNCH(h, (k_mg_vertex_step<double>), grid, 256, 0, (double const*)yv->data, (double*)pv->data,
                  (double const*)ov->data, (double*)xv->data, (int32_t)n_local, alpha, n_vertices_global,
                  first == TRUE ? 1 : 0, totals_prev_device, partial_out_device);
    check_last("pagerank_vertex_step");
  });
}

// cand[row slot] = global code of a frontier source adjacent to that (unvisited) row, or -1.  Asynchronous.
cugraph_error_code_t cugraph_b200_block_bfs_pull(const cugraph_resource_handle_t* handle, cugraph_b200_block_t* block,
                                                 const cugraph_type_erased_device_array_view_t* frontier_cols,
                                                 const cugraph_type_erased_device_array_view_t* visited_rows, size_t maxpart,
                                                 int grid_cols, int grid_c, cugraph_type_erased_device_array_view_t* cand,
                                                 cugraph_error_t** error)
{
  return guarded(error, [&] {
    auto const& h = H(handle);
    B200_EXPECTS(block && frontier_cols && visited_rows && cand, CUGRAPH_INVALID_INPUT, "NULL argument");
    auto* b        = reinterpret_cast<block_impl*>(block);
    auto const* fv = V(frontier_cols);
    auto const* vv = V(visited_rows);
    auto const* cv = V(cand);
    B200_EXPECTS(dtype_size(fv->type) == 1 && dtype_size(vv->type) == 1, CUGRAPH_INVALID_INPUT, "frontier / visited are byte flags");
    B200_EXPECTS(cv->type == INT64, CUGRAPH_INVALID_INPUT, "cand must be INT64");
    B200_EXPECTS(fv->size >= (size_t)b->n_cols && vv->size >= (size_t)b->n_rows && cv->size >= (size_t)b->n_rows,
                 CUGRAPH_INVALID_INPUT, "flag / candidate arrays shorter than the block's slots");
    B200_EXPECTS(maxpart > 0 && grid_cols > 0 && grid_c >= 0 && grid_c < grid_cols, CUGRAPH_INVALID_INPUT, "bad grid position");
    csx_t const& c = *b->csx;
    if (c.offs64)
      block_bfs_pull<int64_t>(h, c, (uint8_t const*)fv->data, (uint8_t const*)vv->data, (long long)maxpart, grid_cols, grid_c,
                              (long long*)cv->data, b->n_rows);
    else
      block_bfs_pull<int32_t>(h, c, (uint8_t const*)fv->data, (uint8_t const*)vv->data, (long long)maxpart, grid_cols, grid_c,
                              (long long*)cv->data, b->n_rows);
    check_last("block_bfs_pull");
  });
}

// the reference's MG constructors take raft comms (not part of this build): see cugraph_b200/mg.py for the torch.distributed path
cugraph_error_code_t cugraph_graph_create_mg(cugraph_resource_handle_t const*, cugraph_graph_properties_t const*,
  cugraph_type_erased_device_array_view_t const* const*, cugraph_type_erased_device_array_view_t const* const*,
  cugraph_type_erased_device_array_view_t const* const*, cugraph_type_erased_device_array_view_t const* const*,
  cugraph_type_erased_device_array_view_t const* const*, cugraph_type_erased_device_array_view_t const* const*,
  bool_t, size_t, bool_t, bool_t, bool_t, bool_t, cugraph_graph_t**, cugraph_error_t** error)
{
  return guarded(error, [&] {
    throw capi_exception(CUGRAPH_NOT_IMPLEMENTED,
                         "cugraph_graph_create_mg needs raft comms; use cugraph_b200.mg.MGGraph (torch.distributed)");
  });
}
cugraph_error_code_t cugraph_graph_create_with_times_mg(cugraph_resource_handle_t const*, cugraph_graph_properties_t const*,
  cugraph_type_erased_device_array_view_t const* const*, cugraph_type_erased_device_array_view_t const* const*,
  cugraph_type_erased_device_array_view_t const* const*, cugraph_type_erased_device_array_view_t const* const*,
  cugraph_type_erased_device_array_view_t const* const*, cugraph_type_erased_device_array_view_t const* const*,
  cugraph_type_erased_device_array_view_t const* const*, cugraph_type_erased_device_array_view_t const* const*,
  bool_t, size_t, bool_t, bool_t, bool_t, bool_t, cugraph_graph_t**, cugraph_error_t** error)
{
  return guarded(error, [&] {
    throw capi_exception(CUGRAPH_NOT_IMPLEMENTED,
                         "cugraph_graph_create_with_times_mg needs raft comms; use cugraph_b200.mg.MGGraph (torch.distributed)");
  });
}

}  // extern "C"
