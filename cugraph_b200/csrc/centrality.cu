// Katz centrality and HITS on the pull sweep — the sibling algorithms that run on the same primitive as PageRank
// (per_v_transform_reduce_incoming_e / _outgoing_e with reduce_op::plus; reference cpp/src/centrality/katz_centrality_impl.cuh:34-196,
// cpp/src/link_analysis/hits_impl.cuh:29-206, C API cpp/src/c_api/katz.cpp, cpp/src/c_api/hits.cpp).  Both are host loops
// over launch_pull_sweep_auto (the shared-memory piece stream when the graph has one) plus small vector passes; their
// per-iteration convergence test reads one scalar back, as the reference does.
#include "sweep.cuh"

#include <cmath>
#include <limits>

namespace b200 {
namespace {

constexpr int kCBlock = 256;
inline int cgrid(handle_impl const& h, int64_t n) { return (int)std::min<int64_t>(std::max<int64_t>((n + kCBlock - 1) / kCBlock, 1), (int64_t)h.sm_count * 8); }

__device__ __forceinline__ double block_sum(double v, double* smem)
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

// out[0] += sum |a - b| ; optionally b <- a (the next sweep's input)
template <typename T>
__global__ void __launch_bounds__(kCBlock) k_abs_diff(T const* __restrict__ a, T* __restrict__ b, int32_t n, int copy, double* __restrict__ out)
{
  __shared__ double smem[kCBlock / 32];
  double d = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    d += fabs((double)a[i] - (double)b[i]);
    if (copy) b[i] = a[i];
  }
  d = block_sum(d, smem);
  if (threadIdx.x == 0 && d != 0.0) atomicAdd(out, d);
}

template <typename T>
__global__ void __launch_bounds__(kCBlock) k_add_vec(T* __restrict__ y, T const* __restrict__ add, int32_t n)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] += add[i];
}

// out[0] += sum v^2 (mode 0) | sum v (mode 1) ; out[1] = max v (mode 2, values are non-negative: integer compare of the bits)
template <typename T>
__global__ void __launch_bounds__(kCBlock) k_norm(T const* __restrict__ v, int32_t n, int mode, double* __restrict__ out)
{
  __shared__ double smem[kCBlock / 32];
  double s = 0.0, m = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double x = (double)v[i];
    s += mode == 0 ? x * x : x;
    m = x > m ? x : m;
  }
  if (mode == 2) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double t = __shfl_xor_sync(0xffffffffu, m, o);
      m              = t > m ? t : m;
    }
    if ((threadIdx.x & 31) == 0 && m > 0.0)
      atomicMax(reinterpret_cast<unsigned long long*>(out + 1), (unsigned long long)__double_as_longlong(m));
    return;
  }
  s = block_sum(s, smem);
  if (threadIdx.x == 0 && s != 0.0) atomicAdd(out, s);
}

template <typename T>
__global__ void __launch_bounds__(kCBlock) k_scale(T* __restrict__ v, int32_t n, double inv)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = (T)((double)v[i] * inv);
}

template <typename T>
__global__ void __launch_bounds__(kCBlock) k_fill_vec(T* __restrict__ v, int64_t n, T val)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = val;
}

template <typename T>
__global__ void k_count_negative(T const* __restrict__ v, int32_t n, int* __restrict__ out)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (v[i] < (T)0) atomicAdd(out, 1);
}

struct sweep_scratch_t {
  dbuf acc, state;
  pr_state_t* st{nullptr};
  void init(handle_impl const& h, size_t rows)
  {
    acc = make_dbuf<double>(std::max<size_t>(rows, 1), h.stream);
    CUDA_TRY(cudaMemsetAsync(acc.data(), 0, sizeof(double) * std::max<size_t>(rows, 1), h.stream));
    state = make_dbuf<pr_state_t>(1, h.stream);
    CUDA_TRY(cudaMemsetAsync(state.data(), 0, sizeof(pr_state_t), h.stream));
    st = state.as<pr_state_t>();
  }
  // the unvarying term the sweep adds to every row
  void set_init(handle_impl const& h, double init)
  {
    pr_state_t hs{};
    hs.init = init;
    CUDA_TRY(cudaMemcpyAsync(state.data(), &hs, sizeof(pr_state_t), cudaMemcpyHostToDevice, h.stream));
    sync(h);  // hs is a stack variable
  }
};

template <typename T>
void sweep(handle_impl const& h, csx_t const& c, int32_t nv, T const* x, T* y, sweep_scratch_t& sc, double alpha, bool use_weights)
{
  if (c.offs64) launch_pull_sweep<int64_t, T>(h, c, x, y, sc.acc.as<double>(), alpha, sc.st, use_weights);
  else launch_pull_sweep_auto<int32_t, T>(h, c, nv, x, y, sc.acc.as<double>(), alpha, sc.st, use_weights);
}

double read_scalar(handle_impl const& h, double const* d)
{
  double v = 0.0;
  CUDA_TRY(cudaMemcpyAsync(&v, d, sizeof(double), cudaMemcpyDeviceToHost, h.stream));
  sync(h);
  return v;
}

// ------------------------------------------------------------------------------------------
// Katz: x <- alpha * A^T x + beta until sum |x_new - x_old| < epsilon; then x / ||x||_2 (the C API always normalises,
// c_api/katz.cpp:116-129, and — faithfully — passes betas = nullptr whatever the caller gave, katz.cpp:151-152)
// ------------------------------------------------------------------------------------------
template <typename T>
void katz_typed(handle_impl const& h, graph_impl& g, double alpha, double beta, double epsilon, size_t max_iterations,
                centrality_result_impl& res)
{
  const int32_t nv = g.n_vertices;
  csx_t const& c   = pull_view(h, g);
  const size_t px  = padded_x_elems(nv, sizeof(T));
  dbuf x = make_dbuf<T>(px, h.stream), y = make_dbuf<T>(std::max(nv, 1), h.stream);
  CUDA_TRY(cudaMemsetAsync(x.data(), 0, px * sizeof(T), h.stream));  // no initial guess: zeros (katz_centrality_impl.cuh:88-93)
  sweep_scratch_t sc;
  sc.init(h, acc_rows(c));
  sc.set_init(h, beta);
  dbuf d_diff = make_dbuf<double>(1, h.stream);
  size_t iter = 0;
  while (nv > 0) {
    sweep<T>(h, c, nv, x.as<T>(), y.as<T>(), sc, alpha, true);
    CUDA_TRY(cudaMemsetAsync(d_diff.data(), 0, sizeof(double), h.stream));
    B200_LAUNCH(h, (k_abs_diff<T>), cgrid(h, nv), kCBlock, 0, y.as<T>(), x.as<T>(), nv, 1, d_diff.as<double>());
    const double diff = read_scalar(h, d_diff.as<double>());
    ++iter;
    if ((T)diff < (T)epsilon) break;
    B200_EXPECTS(iter < max_iterations, CUGRAPH_UNKNOWN_ERROR, "Katz Centrality failed to converge.");
  }
  if (nv > 0) {  // x holds the final values (copied by k_abs_diff)
    CUDA_TRY(cudaMemsetAsync(d_diff.data(), 0, sizeof(double), h.stream));
    B200_LAUNCH(h, (k_norm<T>), cgrid(h, nv), kCBlock, 0, x.as<T>(), nv, 0, d_diff.as<double>());
    const double l2 = std::sqrt(read_scalar(h, d_diff.as<double>()));
    B200_EXPECTS(l2 > 0.0, CUGRAPH_UNKNOWN_ERROR, "L2 norm of the computed Katz Centrality values should be positive.");
    B200_LAUNCH(h, (k_scale<T>), cgrid(h, nv), kCBlock, 0, x.as<T>(), nv, 1.0 / l2);
  }
  res.vertices   = new device_array_impl{reported_vertices(h, g), (size_t)nv, g.vertex_type};
  res.values     = new device_array_impl{to_reported_order(h, g, x.data(), sizeof(T)), (size_t)nv, g.weight_type};
  res.iterations = iter;
  res.converged  = true;
  check_last("katz");
  sync(h);
}

// ------------------------------------------------------------------------------------------
// Eigenvector centrality (eigenvector_centrality_impl.cuh:34-150): x <- (A^T x + x) / ||A^T x + x||_2 from x = 1 / V, until
// sum |x_new - x_old| < V * epsilon
// ------------------------------------------------------------------------------------------
template <typename T>
void eigenvector_typed(handle_impl const& h, graph_impl& g, double epsilon, size_t max_iterations, centrality_result_impl& res)
{
  const int32_t nv = g.n_vertices;
  csx_t const& c   = pull_view(h, g);
  const size_t px  = padded_x_elems(nv, sizeof(T));
  dbuf x = make_dbuf<T>(px, h.stream), y = make_dbuf<T>(std::max(nv, 1), h.stream);
  CUDA_TRY(cudaMemsetAsync(x.data(), 0, px * sizeof(T), h.stream));
  if (nv > 0) B200_LAUNCH(h, (k_fill_vec<T>), cgrid(h, nv), kCBlock, 0, x.as<T>(), (int64_t)nv, (T)(1.0 / (double)nv));
  sweep_scratch_t sc;
  sc.init(h, acc_rows(c));
  dbuf d2     = make_dbuf<double>(2, h.stream);
  size_t iter = 0;
  while (nv > 0) {
    sweep<T>(h, c, nv, x.as<T>(), y.as<T>(), sc, 1.0, true);
    B200_LAUNCH(h, (k_add_vec<T>), cgrid(h, nv), kCBlock, 0, y.as<T>(), x.as<T>(), nv);
    CUDA_TRY(cudaMemsetAsync(d2.data(), 0, 2 * sizeof(double), h.stream));
    B200_LAUNCH(h, (k_norm<T>), cgrid(h, nv), kCBlock, 0, y.as<T>(), nv, 0, d2.as<double>());
    const double hyp = std::sqrt(read_scalar(h, d2.as<double>()));
    B200_LAUNCH(h, (k_scale<T>), cgrid(h, nv), kCBlock, 0, y.as<T>(), nv, 1.0 / hyp);
    CUDA_TRY(cudaMemsetAsync(d2.data(), 0, sizeof(double), h.stream));
    B200_LAUNCH(h, (k_abs_diff<T>), cgrid(h, nv), kCBlock, 0, y.as<T>(), x.as<T>(), nv, 1, d2.as<double>());
    const double diff = read_scalar(h, d2.as<double>());
    ++iter;
    if ((T)diff < (T)nv * (T)epsilon) break;
    B200_EXPECTS(iter < max_iterations, CUGRAPH_UNKNOWN_ERROR, "Eigenvector Centrality failed to converge.");
  }
  res.vertices   = new device_array_impl{reported_vertices(h, g), (size_t)nv, g.vertex_type};
  res.values     = new device_array_impl{to_reported_order(h, g, x.data(), sizeof(T)), (size_t)nv, g.weight_type};
  res.iterations = iter;
  res.converged  = true;
  check_last("eigenvector_centrality");
  sync(h);
}

// ------------------------------------------------------------------------------------------
// HITS (hits_impl.cuh:49-191): authorities = sum over in-edges of the hubs, hubs = sum over out-edges of the
// authorities, both divided by their maximum; until sum |hubs - previous hubs| < V * epsilon; edge weights are not used
// ------------------------------------------------------------------------------------------
struct hits_result_impl {
  device_array_impl* vertices{nullptr};
  device_array_impl* hubs{nullptr};
  device_array_impl* authorities{nullptr};
  double hub_score_differences{0.0};
  size_t number_of_iterations{0};
};

template <typename T>
void normalize_by(handle_impl const& h, T* v, int32_t nv, int mode, dbuf& d2)
{
  CUDA_TRY(cudaMemsetAsync(d2.data(), 0, 2 * sizeof(double), h.stream));
  B200_LAUNCH(h, (k_norm<T>), cgrid(h, nv), kCBlock, 0, v, nv, mode, d2.as<double>());
  const double norm = read_scalar(h, d2.as<double>() + (mode == 2 ? 1 : 0));
  B200_EXPECTS((T)norm > (T)0, CUGRAPH_UNKNOWN_ERROR, "Norm is required to be a positive value.");
  B200_LAUNCH(h, (k_scale<T>), cgrid(h, nv), kCBlock, 0, v, nv, 1.0 / norm);
}

template <typename T>
void hits_typed(handle_impl const& h, graph_impl& g, double epsilon, size_t max_iterations, device_array_view_impl const* guess_v,
                device_array_view_impl const* guess_val, bool normalize, bool do_expensive_check, hits_result_impl& res)
{
  const int32_t nv   = g.n_vertices;
  csx_t const& c_in  = pull_view(h, g);       // rows = destinations: authorities <- hubs
  csx_t const& c_out = out_sweep_view(h, g);  // rows = sources: hubs <- authorities
  const size_t px    = padded_x_elems(nv, sizeof(T));
  dbuf hubs_a = make_dbuf<T>(px, h.stream), hubs_b = make_dbuf<T>(px, h.stream), auth = make_dbuf<T>(px, h.stream);
  for (dbuf* b : {&hubs_a, &hubs_b, &auth}) CUDA_TRY(cudaMemsetAsync(b->data(), 0, px * sizeof(T), h.stream));
  dbuf d2 = make_dbuf<double>(2, h.stream);
  B200_EXPECTS(epsilon >= 0.0, CUGRAPH_INVALID_INPUT, "Invalid input argument: epsilon should be non-negative.");
  double diff = std::numeric_limits<T>::max();
  size_t iter = max_iterations;
  if (nv > 0) {
    const T tolerance = (T)nv * (T)epsilon;
    if (guess_v) {
      dbuf gv = collect_vertex_values<T>(h, g, guess_v, guess_val, (T)0);
      CUDA_TRY(cudaMemcpyAsync(hubs_a.data(), gv.data(), sizeof(T) * nv, cudaMemcpyDeviceToDevice, h.stream));
      if (do_expensive_check) {
        dbuf neg = make_dbuf<int>(1, h.stream);
        CUDA_TRY(cudaMemsetAsync(neg.data(), 0, sizeof(int), h.stream));
        B200_LAUNCH(h, (k_count_negative<T>), cgrid(h, nv), kCBlock, 0, hubs_a.as<T>(), nv, neg.as<int>());
        int hn = 0;
        CUDA_TRY(cudaMemcpyAsync(&hn, neg.data(), sizeof(int), cudaMemcpyDeviceToHost, h.stream));
        sync(h);
        B200_EXPECTS(hn == 0, CUGRAPH_INVALID_INPUT, "Invalid input argument: initial guess values should be non-negative.");
      }
      normalize_by<T>(h, hubs_a.as<T>(), nv, 1, d2);
    } else {
      B200_LAUNCH(h, (k_fill_vec<T>), cgrid(h, nv), kCBlock, 0, hubs_a.as<T>(), (int64_t)nv, (T)(1.0 / (double)nv));
    }
    sweep_scratch_t sc_in, sc_out;
    sc_in.init(h, acc_rows(c_in));
    sc_out.init(h, acc_rows(c_out));
    T* prev = hubs_a.as<T>();
    T* curr = hubs_b.as<T>();
    iter    = 0;
    while (true) {
      sweep<T>(h, c_in, nv, prev, auth.as<T>(), sc_in, 1.0, false);
      sweep<T>(h, c_out, nv, auth.as<T>(), curr, sc_out, 1.0, false);
      normalize_by<T>(h, curr, nv, 2, d2);
      normalize_by<T>(h, auth.as<T>(), nv, 2, d2);
      CUDA_TRY(cudaMemsetAsync(d2.data(), 0, sizeof(double), h.stream));
      B200_LAUNCH(h, (k_abs_diff<T>), cgrid(h, nv), kCBlock, 0, curr, prev, nv, 0, d2.as<double>());
      diff = (double)(T)read_scalar(h, d2.as<double>());
      std::swap(prev, curr);
      ++iter;
      if ((T)diff < tolerance) break;
      B200_EXPECTS(iter < max_iterations, CUGRAPH_UNKNOWN_ERROR, "HITS failed to converge.");
    }
    if (normalize) {
      normalize_by<T>(h, prev, nv, 1, d2);
      normalize_by<T>(h, auth.as<T>(), nv, 1, d2);
    }
    res.hubs        = new device_array_impl{to_reported_order(h, g, prev, sizeof(T)), (size_t)nv, g.weight_type};
    res.authorities = new device_array_impl{to_reported_order(h, g, auth.data(), sizeof(T)), (size_t)nv, g.weight_type};
  } else {
    res.hubs        = new device_array_impl{dbuf(0, h.stream), 0, g.weight_type};
    res.authorities = new device_array_impl{dbuf(0, h.stream), 0, g.weight_type};
  }
  res.vertices              = new device_array_impl{reported_vertices(h, g), (size_t)nv, g.vertex_type};
  res.hub_score_differences = diff;
  res.number_of_iterations  = iter;
  check_last("hits");
  sync(h);
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

cugraph_error_code_t cugraph_katz_centrality(const cugraph_resource_handle_t* handle, cugraph_graph_t* graph,
                                             const cugraph_type_erased_device_array_view_t* betas, double alpha, double beta,
                                             double epsilon, size_t max_iterations, bool_t do_expensive_check,
                                             cugraph_centrality_result_t** result, cugraph_error_t** error)
{
  (void)betas;  // the reference's C entry point drops them (c_api/katz.cpp:151-152 constructs its functor with nullptr)
  (void)do_expensive_check;
  return guarded(error, [&] {
    auto const& h = H(handle);
    auto* g       = G(graph);
    B200_EXPECTS(result != nullptr, CUGRAPH_INVALID_INPUT, "result out-pointer is NULL");
    *result = nullptr;
    B200_EXPECTS(g->mg == nullptr, CUGRAPH_NOT_IMPLEMENTED, "multi-GPU Katz centrality is not implemented");
    B200_EXPECTS(alpha >= 0.0 && alpha <= 1.0, CUGRAPH_INVALID_INPUT, "Invalid input argument: alpha should be in [0.0, 1.0].");
    B200_EXPECTS(epsilon >= 0.0, CUGRAPH_INVALID_INPUT, "Invalid input argument: epsilon should be non-negative.");
    auto res = std::make_unique<centrality_result_impl>();
    if (g->weight_type == FLOAT32) katz_typed<float>(h, *g, alpha, beta, epsilon, max_iterations, *res);
    else katz_typed<double>(h, *g, alpha, beta, epsilon, max_iterations, *res);
    *result = reinterpret_cast<cugraph_centrality_result_t*>(res.release());
  });
}

cugraph_error_code_t cugraph_eigenvector_centrality(const cugraph_resource_handle_t* handle, cugraph_graph_t* graph, double epsilon,
                                                    size_t max_iterations, bool_t do_expensive_check,
                                                    cugraph_centrality_result_t** result, cugraph_error_t** error)
{
  (void)do_expensive_check;
  return guarded(error, [&] {
    auto const& h = H(handle);
    auto* g       = G(graph);
    B200_EXPECTS(result != nullptr, CUGRAPH_INVALID_INPUT, "result out-pointer is NULL");
    *result = nullptr;
    B200_EXPECTS(g->mg == nullptr, CUGRAPH_NOT_IMPLEMENTED, "multi-GPU eigenvector centrality is not implemented");
    B200_EXPECTS(epsilon >= 0.0, CUGRAPH_INVALID_INPUT, "Invalid input argument: epsilon should be non-negative.");
    auto res = std::make_unique<centrality_result_impl>();
    if (g->weight_type == FLOAT32) eigenvector_typed<float>(h, *g, epsilon, max_iterations, *res);
    else eigenvector_typed<double>(h, *g, epsilon, max_iterations, *res);
    *result = reinterpret_cast<cugraph_centrality_result_t*>(res.release());
  });
}

cugraph_error_code_t cugraph_hits(const cugraph_resource_handle_t* handle, cugraph_graph_t* graph, double epsilon,
                                  size_t max_iterations, const cugraph_type_erased_device_array_view_t* initial_hubs_guess_vertices,
                                  const cugraph_type_erased_device_array_view_t* initial_hubs_guess_values, bool_t normalize,
                                  bool_t do_expensive_check, cugraph_hits_result_t** result, cugraph_error_t** error)
{
  return guarded(error, [&] {
    auto const& h = H(handle);
    auto* g       = G(graph);
    B200_EXPECTS(result != nullptr, CUGRAPH_INVALID_INPUT, "result out-pointer is NULL");
    *result = nullptr;
    B200_EXPECTS(g->mg == nullptr, CUGRAPH_NOT_IMPLEMENTED, "multi-GPU HITS is not implemented");
    auto const* gv = V(initial_hubs_guess_vertices);
    auto const* gx = V(initial_hubs_guess_values);
    if (gv) {
      B200_EXPECTS(gx != nullptr && gx->size == gv->size, CUGRAPH_INVALID_INPUT, "initial hubs guess needs vertices and values of equal size");
      B200_EXPECTS(gv->type == g->vertex_type, CUGRAPH_INVALID_INPUT, "vertex type of graph and initial_hubs_guess_vertices must match");
      B200_EXPECTS(gx->type == g->weight_type, CUGRAPH_INVALID_INPUT, "weight type of graph and initial_hubs_guess_values must match");
    }
    auto res = std::make_unique<hits_result_impl>();
    if (g->weight_type == FLOAT32)
      hits_typed<float>(h, *g, epsilon, max_iterations, gv, gx, normalize == TRUE, do_expensive_check == TRUE, *res);
    else
      hits_typed<double>(h, *g, epsilon, max_iterations, gv, gx, normalize == TRUE, do_expensive_check == TRUE, *res);
    *result = reinterpret_cast<cugraph_hits_result_t*>(res.release());
  });
}

cugraph_type_erased_device_array_view_t* cugraph_hits_result_get_vertices(cugraph_hits_result_t* result)
{
  return reinterpret_cast<cugraph_type_erased_device_array_view_t*>(reinterpret_cast<hits_result_impl*>(result)->vertices->new_view());
}
cugraph_type_erased_device_array_view_t* cugraph_hits_result_get_hubs(cugraph_hits_result_t* result)
{
  return reinterpret_cast<cugraph_type_erased_device_array_view_t*>(reinterpret_cast<hits_result_impl*>(result)->hubs->new_view());
}
cugraph_type_erased_device_array_view_t* cugraph_hits_result_get_authorities(cugraph_hits_result_t* result)
{
  return reinterpret_cast<cugraph_type_erased_device_array_view_t*>(reinterpret_cast<hits_result_impl*>(result)->authorities->new_view());
}
double cugraph_hits_result_get_hub_score_differences(cugraph_hits_result_t* result)
{
  return reinterpret_cast<hits_result_impl*>(result)->hub_score_differences;
}
size_t cugraph_hits_result_get_number_of_iterations(cugraph_hits_result_t* result)
{
  return reinterpret_cast<hits_result_impl*>(result)->number_of_iterations;
}
void cugraph_hits_result_free(cugraph_hits_result_t* result)
{
  if (!result) return;
  auto* r = reinterpret_cast<hits_result_impl*>(result);
  delete r->vertices;
  delete r->hubs;
  delete r->authorities;
  delete r;
}

}  // extern "C"
