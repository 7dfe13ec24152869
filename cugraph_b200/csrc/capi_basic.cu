// C-ABI entry points that carry no graph logic: errors, resource handle, type-erased arrays.
// Boundary being replaced: cpp/src/c_api/{error,resource_handle,array}.cpp of the reference.
#include "graph.cuh"

#include <cstdlib>
#include <map>
#include <mutex>
#include <unordered_map>
#include <unordered_set>

namespace b200 {
namespace {
std::mutex g_stream_mutex;
std::unordered_set<cudaStream_t>& live_streams()
{
  static auto* s = new std::unordered_set<cudaStream_t>();  // leaked on purpose: used during exit
  return *s;
}
// Freed blocks per stream (see block_alloc in common.cuh).  Leaked on purpose like the stream set.
struct block_cache_t {
  std::multimap<size_t, void*> blocks;  // capacity -> block
  size_t total{0};
};
std::unordered_map<cudaStream_t, block_cache_t>& block_caches()
{
  static auto* c = new std::unordered_map<cudaStream_t, block_cache_t>();
  return *c;
}
constexpr size_t kBlockCacheCap = 32ull << 30;  // bytes kept per stream; beyond it the cache of that stream is returned to the pool

// caller holds g_stream_mutex
void flush_cache_locked(cudaStream_t s, block_cache_t& c, bool stream_alive)
{
  for (auto& b : c.blocks) {
    if (stream_alive) cudaFreeAsync(b.second, s);
    else cudaFree(b.second);
  }
  c.blocks.clear();
  c.total = 0;
}
}  // namespace

void* block_alloc(size_t bytes, cudaStream_t s, size_t* capacity)
{
  {
    std::lock_guard<std::mutex> lk(g_stream_mutex);
    auto it = block_caches().find(s);
    if (it != block_caches().end()) {
      auto b = it->second.blocks.lower_bound(bytes);
      if (b != it->second.blocks.end() && b->first <= bytes + bytes / 8 + 512) {
        void* p   = b->second;
        *capacity = b->first;
        it->second.total -= b->first;
        it->second.blocks.erase(b);
        return p;
      }
    }
  }
  void* p       = nullptr;
  cudaError_t e = cudaMallocAsync(&p, bytes, s);
  if (e != cudaSuccess) {  // give everything cached back and try once more
    (void)cudaGetLastError();
    {
      std::lock_guard<std::mutex> lk(g_stream_mutex);
      for (auto& kv : block_caches()) flush_cache_locked(kv.first, kv.second, live_streams().count(kv.first) != 0);
    }
    cudaDeviceSynchronize();
    e = cudaMallocAsync(&p, bytes, s);
  }
  CUDA_TRY(e);
  *capacity = bytes;
  return p;
}

void block_free(void* p, size_t capacity, cudaStream_t s)
{
  std::lock_guard<std::mutex> lk(g_stream_mutex);
  if (live_streams().count(s) == 0) {
    cudaFree(p);
    return;
  }
  block_cache_t& c = block_caches()[s];
  if (c.total + capacity > kBlockCacheCap) flush_cache_locked(s, c, true);
  if (capacity > kBlockCacheCap) {
    cudaFreeAsync(p, s);
    return;
  }
  c.blocks.emplace(capacity, p);
  c.total += capacity;
}

bool stream_is_live(cudaStream_t s)
{
  std::lock_guard<std::mutex> lk(g_stream_mutex);
  return live_streams().count(s) != 0;
}
void register_stream(cudaStream_t s)
{
  std::lock_guard<std::mutex> lk(g_stream_mutex);
  live_streams().insert(s);
}
void unregister_stream(cudaStream_t s)
{
  std::lock_guard<std::mutex> lk(g_stream_mutex);
  auto it = block_caches().find(s);
  if (it != block_caches().end()) {  // the caller synchronised the stream: the cached blocks are idle
    flush_cache_locked(s, it->second, true);
    block_caches().erase(it);
  }
  live_streams().erase(s);
}
}  // namespace b200

using namespace b200;

extern "C" {

// ----------------------------------------------------------------------------- error.h:28-29
const char* cugraph_error_message(const cugraph_error_t* error)
{
  if (error == nullptr) return nullptr;
  return reinterpret_cast<error_impl const*>(error)->message.c_str();
}

void cugraph_error_free(cugraph_error_t* error)
{
  if (error != nullptr) delete reinterpret_cast<error_impl*>(error);
}

// ------------------------------------------------------------------- resource_handle.h:25-31
// NULL -> single-GPU handle on the current device.  Non-NULL (a raft handle in the reference) is refused (mg.cu: attach_comm).
cugraph_resource_handle_t* cugraph_create_resource_handle(void* raft_handle)
{
  try {
    auto* h = new handle_impl{};
    h->tune = tuning_t::from_env();
    CUDA_TRY(cudaGetDevice(&h->device));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
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
    // keep freed blocks in the pool: algorithm calls allocate/free V- and E-sized scratch
    cudaMemPool_t pool;
    CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, h->device));
    uint64_t threshold = UINT64_MAX;
    CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    if (raft_handle != nullptr) {
      attach_comm(h, raft_handle);
    }
    return reinterpret_cast<cugraph_resource_handle_t*>(h);
  } catch (std::exception const& e) {
    std::fprintf(stderr, "cugraph_create_resource_handle: %s\n", e.what());
    return nullptr;
  }
}

int cugraph_resource_handle_get_comm_size(const cugraph_resource_handle_t* handle)
{
  return handle ? reinterpret_cast<handle_impl const*>(handle)->size : 1;
}

int cugraph_resource_handle_get_rank(const cugraph_resource_handle_t* handle)
{
  return handle ? reinterpret_cast<handle_impl const*>(handle)->rank : 0;
}

void cugraph_free_resource_handle(cugraph_resource_handle_t* handle)
{
  if (!handle) return;
  auto* h = reinterpret_cast<handle_impl*>(handle);
  cudaStreamSynchronize(h->stream);
  cudaStreamSynchronize(h->aux_stream);
  unregister_stream(h->stream);
  unregister_stream(h->aux_stream);
  cudaEventDestroy(h->ev_a);
  cudaEventDestroy(h->ev_b);
  if (!h->borrowed_stream) cudaStreamDestroy(h->stream);
  cudaStreamDestroy(h->aux_stream);
  cudaFreeHost(h->pinned);
  delete h;
}

// ------------------------------------------------------------------------------ b200_ext.h
const char* cugraph_b200_version(void) { return "cugraph_b200 0.1 (sm_100a)"; }

void* cugraph_b200_handle_stream(const cugraph_resource_handle_t* handle)
{
  return handle ? reinterpret_cast<void*>(reinterpret_cast<handle_impl const*>(handle)->stream) : nullptr;
}

size_t cugraph_b200_handle_launch_count(const cugraph_resource_handle_t* handle)
{
  return handle ? reinterpret_cast<handle_impl const*>(handle)->launches : 0;
}

// -------------------------------------------------------------------------- array.h:43-121
cugraph_error_code_t cugraph_type_erased_device_array_create(const cugraph_resource_handle_t* handle,
                                                             size_t n_elems,
                                                             cugraph_data_type_id_t dtype,
                                                             cugraph_type_erased_device_array_t** array,
                                                             cugraph_error_t** error)
{
  return guarded(error, [&] {
    auto const& h = H(handle);
    B200_EXPECTS(array != nullptr, CUGRAPH_INVALID_INPUT, "array out-pointer is NULL");
    size_t es = dtype_size(dtype);
    B200_EXPECTS(es > 0, CUGRAPH_UNSUPPORTED_TYPE_COMBINATION, "invalid dtype");
    dbuf b(n_elems * es, h.stream);
    *array = wrap_array(std::move(b), n_elems, dtype);
  });
}

cugraph_error_code_t cugraph_type_erased_device_array_create_from_view(
  const cugraph_resource_handle_t* handle,
  const cugraph_type_erased_device_array_view_t* view,
  cugraph_type_erased_device_array_t** array,
  cugraph_error_t** error)
{
  return guarded(error, [&] {
    auto const& h = H(handle);
    B200_EXPECTS(view != nullptr && array != nullptr, CUGRAPH_INVALID_INPUT, "NULL argument");
    auto const* v = V(view);
    dbuf b(v->nbytes(), h.stream);
    if (v->nbytes() > 0)
      CUDA_TRY(cudaMemcpyAsync(b.data(), v->data, v->nbytes(), cudaMemcpyDeviceToDevice, h.stream));
    sync(h);
    *array = wrap_array(std::move(b), v->size, v->type);
  });
}

void cugraph_type_erased_device_array_free(cugraph_type_erased_device_array_t* p)
{
  if (p) delete reinterpret_cast<device_array_impl*>(p);
}

cugraph_type_erased_device_array_view_t* cugraph_type_erased_device_array_view(
  cugraph_type_erased_device_array_t* array)
{
  if (!array) return nullptr;
  return reinterpret_cast<cugraph_type_erased_device_array_view_t*>(
    reinterpret_cast<device_array_impl*>(array)->new_view());
}

cugraph_error_code_t cugraph_type_erased_device_array_view_as_type(
  cugraph_type_erased_device_array_t* array,
  cugraph_data_type_id_t dtype,
  cugraph_type_erased_device_array_view_t** result_view,
  cugraph_error_t** error)
{
  // reinterpretation is only allowed between types of equal width (reference array.cpp)
  return guarded(error, [&] {
    B200_EXPECTS(array != nullptr && result_view != nullptr, CUGRAPH_INVALID_INPUT, "NULL argument");
    auto* a = reinterpret_cast<device_array_impl*>(array);
    B200_EXPECTS(dtype_size(dtype) == dtype_size(a->type) && dtype_size(dtype) > 0,
                 CUGRAPH_INVALID_INPUT,
                 "Could not treat type_erased_device_array_t as requested type");
    *result_view = reinterpret_cast<cugraph_type_erased_device_array_view_t*>(
      new device_array_view_impl{a->buf.data(), a->size, dtype});
  });
}

// ------------------------------------------------------------------------- array.h:123-170
cugraph_type_erased_device_array_view_t* cugraph_type_erased_device_array_view_create(
  void* pointer, size_t n_elems, cugraph_data_type_id_t dtype)
{
  return reinterpret_cast<cugraph_type_erased_device_array_view_t*>(
    new device_array_view_impl{pointer, n_elems, dtype});
}

void cugraph_type_erased_device_array_view_free(cugraph_type_erased_device_array_view_t* p)
{
  if (p) delete reinterpret_cast<device_array_view_impl*>(p);
}

size_t cugraph_type_erased_device_array_view_size(const cugraph_type_erased_device_array_view_t* p)
{
  return p ? V(p)->size : 0;
}

cugraph_data_type_id_t cugraph_type_erased_device_array_view_type(
  const cugraph_type_erased_device_array_view_t* p)
{
  return p ? V(p)->type : NTYPES;
}

const void* cugraph_type_erased_device_array_view_pointer(const cugraph_type_erased_device_array_view_t* p)
{
  return p ? V(p)->data : nullptr;
}

// ------------------------------------------------------------------------- array.h:172-262
cugraph_error_code_t cugraph_type_erased_host_array_create(const cugraph_resource_handle_t* handle,
                                                           size_t n_elems,
                                                           cugraph_data_type_id_t dtype,
                                                           cugraph_type_erased_host_array_t** array,
                                                           cugraph_error_t** error)
{
  return guarded(error, [&] {
    (void)H(handle);
    B200_EXPECTS(array != nullptr, CUGRAPH_INVALID_INPUT, "array out-pointer is NULL");
    size_t es = dtype_size(dtype);
    B200_EXPECTS(es > 0, CUGRAPH_UNSUPPORTED_TYPE_COMBINATION, "invalid dtype");
    void* p = std::malloc(n_elems * es > 0 ? n_elems * es : 1);
    B200_EXPECTS(p != nullptr, CUGRAPH_ALLOC_ERROR, "host allocation failed");
    *array = reinterpret_cast<cugraph_type_erased_host_array_t*>(new host_array_impl{p, n_elems, dtype});
  });
}

void cugraph_type_erased_host_array_free(cugraph_type_erased_host_array_t* p)
{
  if (!p) return;
  auto* a = reinterpret_cast<host_array_impl*>(p);
  std::free(a->data);
  delete a;
}

cugraph_type_erased_host_array_view_t* cugraph_type_erased_host_array_view(
  cugraph_type_erased_host_array_t* array)
{
  if (!array) return nullptr;
  auto* a = reinterpret_cast<host_array_impl*>(array);
  return reinterpret_cast<cugraph_type_erased_host_array_view_t*>(
    new host_array_view_impl{a->data, a->size, a->type});
}

cugraph_type_erased_host_array_view_t* cugraph_type_erased_host_array_view_create(
  void* pointer, size_t n_elems, cugraph_data_type_id_t dtype)
{
  return reinterpret_cast<cugraph_type_erased_host_array_view_t*>(
    new host_array_view_impl{pointer, n_elems, dtype});
}

void cugraph_type_erased_host_array_view_free(cugraph_type_erased_host_array_view_t* p)
{
  if (p) delete reinterpret_cast<host_array_view_impl*>(p);
}

size_t cugraph_type_erased_host_array_size(const cugraph_type_erased_host_array_view_t* p)
{
  return p ? reinterpret_cast<host_array_view_impl const*>(p)->size : 0;
}

cugraph_data_type_id_t cugraph_type_erased_host_array_type(const cugraph_type_erased_host_array_view_t* p)
{
  return p ? reinterpret_cast<host_array_view_impl const*>(p)->type : NTYPES;
}

void* cugraph_type_erased_host_array_pointer(const cugraph_type_erased_host_array_view_t* p)
{
  return p ? reinterpret_cast<host_array_view_impl const*>(p)->data : nullptr;
}

// ------------------------------------------------------------------------- array.h:264-326
cugraph_error_code_t cugraph_type_erased_host_array_view_copy(
  const cugraph_resource_handle_t* handle,
  cugraph_type_erased_host_array_view_t* dst,
  const cugraph_type_erased_host_array_view_t* src,
  cugraph_error_t** error)
{
  return guarded(error, [&] {
    (void)H(handle);
    B200_EXPECTS(dst && src, CUGRAPH_INVALID_INPUT, "NULL argument");
    auto* d       = reinterpret_cast<host_array_view_impl*>(dst);
    auto const* s = reinterpret_cast<host_array_view_impl const*>(src);
    B200_EXPECTS(d->nbytes() == s->nbytes(), CUGRAPH_INVALID_INPUT,
                 "source and destination arrays are different sizes");
    std::memcpy(d->data, s->data, s->nbytes());
  });
}

cugraph_error_code_t cugraph_type_erased_device_array_view_copy_from_host(
  const cugraph_resource_handle_t* handle,
  cugraph_type_erased_device_array_view_t* dst,
  const byte_t* h_src,
  cugraph_error_t** error)
{
  return guarded(error, [&] {
    auto const& h = H(handle);
    B200_EXPECTS(dst != nullptr, CUGRAPH_INVALID_INPUT, "NULL argument");
    auto* d = reinterpret_cast<device_array_view_impl*>(dst);
    if (d->nbytes() > 0) {
      B200_EXPECTS(h_src != nullptr, CUGRAPH_INVALID_INPUT, "host source is NULL");
      CUDA_TRY(cudaMemcpyAsync(d->data, h_src, d->nbytes(), cudaMemcpyHostToDevice, h.stream));
    }
    sync(h);
  });
}

cugraph_error_code_t cugraph_type_erased_device_array_view_copy_to_host(
  const cugraph_resource_handle_t* handle,
  byte_t* h_dst,
  const cugraph_type_erased_device_array_view_t* src,
  cugraph_error_t** error)
{
  return guarded(error, [&] {
    auto const& h = H(handle);
    B200_EXPECTS(src != nullptr, CUGRAPH_INVALID_INPUT, "NULL argument");
    auto const* s = V(src);
    if (s->nbytes() > 0) {
      B200_EXPECTS(h_dst != nullptr, CUGRAPH_INVALID_INPUT, "host destination is NULL");
      CUDA_TRY(cudaMemcpyAsync(h_dst, s->data, s->nbytes(), cudaMemcpyDeviceToHost, h.stream));
    }
    sync(h);
  });
}

cugraph_error_code_t cugraph_type_erased_device_array_view_copy(
  const cugraph_resource_handle_t* handle,
  cugraph_type_erased_device_array_view_t* dst,
  const cugraph_type_erased_device_array_view_t* src,
  cugraph_error_t** error)
{
  return guarded(error, [&] {
    auto const& h = H(handle);
    B200_EXPECTS(dst && src, CUGRAPH_INVALID_INPUT, "NULL argument");
    auto* d       = reinterpret_cast<device_array_view_impl*>(dst);
    auto const* s = V(src);
    B200_EXPECTS(d->nbytes() == s->nbytes(), CUGRAPH_INVALID_INPUT,
                 "source and destination arrays are different sizes");
    if (s->nbytes() > 0)
      CUDA_TRY(cudaMemcpyAsync(d->data, s->data, s->nbytes(), cudaMemcpyDeviceToDevice, h.stream));
    sync(h);
  });
}

}  // extern "C"
