// Development probe WITHOUT a Python start-up: on a fresh GPU box `import torch` alone can cost a minute of the GPU budget,
// a C program against the C ABI starts at once (the round's last GPU call ran three C test programs in 4.5 s).
// RMAT edge list generated on the device (counter-based hash RNG, Graph500 a/b/c; statistically the bench graph, not the
// same draws as cugraph_b200/generators.py), graph through the C ABI, then per mode one JSON line:
//   cbench <scale> sweep      parity of the configured sweep vs the plain one + sweep time (best of 3 x 20)
//   cbench <scale> pagerank   two calls of 100 iterations, the second timed
//   cbench <scale> trav [n]   BFS (direction-optimising) and SSSP from n sources on the symmetrised weighted graph
//   cbench <scale> all [n]
// Environment switches (CUGRAPH_B200_*) apply as everywhere.  bench.py stays the measurement of record.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -I include scripts/cbench.cu -o cugraph_b200/lib/cbench \
//        -L cugraph_b200/lib -l:libcugraph_c.so -Xlinker -rpath -Xlinker '$ORIGIN'
#include <cugraph_c/algorithms.h>
#include <cugraph_c/b200_ext.h>
#include <cugraph_c/graph.h>

#include <cuda_runtime.h>  // with -DB200_HOST_EMU -I emu: the CPU emulation shim (logic check of this program without a GPU)

#include <chrono>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(2); } \
  } while (0)
#define CG(x)                                                                                         \
  do {                                                                                                \
    cugraph_error_code_t c_ = (x);                                                                    \
    if (c_ != CUGRAPH_SUCCESS) { std::fprintf(stderr, "%s: %s\n", #x, err ? cugraph_error_message(err) : "?"); std::exit(3); } \
  } while (0)

__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__global__ void k_rmat(int scale, long long n, unsigned long long seed, int32_t* src, int32_t* dst, float* w)
{
  const float ab = 0.57f + 0.19f, a_norm = 0.57f / ab, c_norm = 0.19f / (1.0f - ab);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int s = 0, d = 0;
    for (int bit = scale - 1; bit >= 0; --bit) {
      const unsigned long long r = mix64(seed ^ ((unsigned long long)e * 64ull + (unsigned)bit));
      const float r0 = (float)(r >> 40) * (1.0f / 16777216.0f), r1 = (float)((r >> 8) & 0xffffffu) * (1.0f / 16777216.0f);
      const int sb = r0 > ab;
      const int db = r1 > (sb ? c_norm : a_norm);
      s |= sb << bit;
      d |= db << bit;
    }
    src[e] = s;
    dst[e] = d;
    if (w) w[e] = (float)(mix64(seed + 77 + (unsigned long long)e) >> 40) * (1.0f / 16777216.0f);
  }
}

__global__ void k_mirror(long long n, int32_t* src, int32_t* dst, float* w)
{
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    src[n + e] = dst[e];
    dst[n + e] = src[e];
    if (w) w[n + e] = w[e];
  }
}

#ifdef B200_HOST_EMU
#define LAUNCH(kernel, grid, block, ...) emu_launch((grid), (block), [&] { kernel(__VA_ARGS__); })
#else
#define LAUNCH(kernel, grid, block, ...) kernel<<<(grid), (block)>>>(__VA_ARGS__)
#endif

static double now_ms()
{
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static cugraph_type_erased_device_array_view_t* view(void* p, size_t n, cugraph_data_type_id_t t)
{
  return cugraph_type_erased_device_array_view_create(p, n, t);
}

int main(int argc, char** argv)
{
  if (argc < 3) { std::fprintf(stderr, "usage: cbench <scale> sweep|pagerank|trav|all [sources]\n"); return 1; }
  const int scale        = std::atoi(argv[1]);
  const char* mode       = argv[2];
  const int n_src        = argc > 3 ? std::atoi(argv[3]) : 8;
  const bool all         = !std::strcmp(mode, "all");
  const long long E      = 16ll << scale;
  cugraph_error_t* err   = nullptr;
  cugraph_resource_handle_t* h = cugraph_create_resource_handle(nullptr);
  if (!h) { std::fprintf(stderr, "no handle\n"); return 2; }
  int32_t *src, *dst;
  float* w;
  CK(cudaMalloc(&src, sizeof(int32_t) * 2 * E));
  CK(cudaMalloc(&dst, sizeof(int32_t) * 2 * E));
  CK(cudaMalloc(&w, sizeof(float) * 2 * E));
  LAUNCH(k_rmat, 148 * 16, 256, scale, E, 12345ull, src, dst, w);
  CK(cudaDeviceSynchronize());
  cugraph_graph_properties_t props;
  props.is_symmetric  = FALSE;
  props.is_multigraph = TRUE;

  if (all || !std::strcmp(mode, "sweep") || !std::strcmp(mode, "pagerank")) {
    cugraph_graph_t* g = nullptr;
    auto *vs = view(src, E, INT32), *vd = view(dst, E, INT32);
    double t0 = now_ms();
    CG(cugraph_graph_create_with_times_sg(h, &props, nullptr, vs, vd, nullptr, nullptr, nullptr, nullptr, nullptr, TRUE, TRUE,
                                          FALSE, FALSE, FALSE, FALSE, &g, &err));
    const double create_ms = now_ms() - t0;
    if (all || !std::strcmp(mode, "sweep")) {
      double cmp[8];
      CG(cugraph_b200_debug_compare_sweeps(h, g, cmp, &err));
      double ms = 0, by = 0, best = 1e30;
      for (int r = 0; r < 3; ++r) {
        CG(cugraph_b200_time_pull_spmv(h, g, 20, &ms, &by, &err));
        if (ms < best) best = ms;
      }
      std::printf("{\"mode\": \"sweep\", \"scale\": %d, \"create_ms\": %.2f, \"sweep_ms\": %.4f, \"sweep_gbs\": %.1f, "
                  "\"max_rel_diff_ge32\": %.3e, \"bad_rows_ge32\": %.0f, \"max_rel_diff_lt32\": %.3e, \"bad_rows_lt32\": %.0f}\n",
                  scale, create_ms, best, by / best / 1e6, cmp[0], cmp[3], cmp[4], cmp[7]);
    }
    if (all || !std::strcmp(mode, "pagerank")) {
      double t_call = 0;
      for (int r = 0; r < 2; ++r) {
        cugraph_centrality_result_t* res = nullptr;
        t0 = now_ms();
        CG(cugraph_pagerank_allow_nonconvergence(h, g, nullptr, nullptr, nullptr, nullptr, 0.85, 0.0, 100, FALSE, &res, &err));
        t_call = now_ms() - t0;
        cugraph_centrality_result_free(res);
      }
      std::printf("{\"mode\": \"pagerank\", \"scale\": %d, \"pagerank100_ms\": %.3f, \"mteps\": %.0f, \"launches\": %zu}\n", scale,
                  t_call, (double)E * 100.0 / t_call / 1e3, cugraph_b200_handle_launch_count(h));
    }
    cugraph_graph_free(g);
    cugraph_type_erased_device_array_view_free(vs);
    cugraph_type_erased_device_array_view_free(vd);
  }

  if (all || !std::strcmp(mode, "trav")) {
    LAUNCH(k_mirror, 148 * 16, 256, E, src, dst, w);
    CK(cudaDeviceSynchronize());
    props.is_symmetric = TRUE;
    cugraph_graph_t* g = nullptr;
    auto *vs = view(src, 2 * E, INT32), *vd = view(dst, 2 * E, INT32), *vw = view(w, 2 * E, FLOAT32);
    double t0 = now_ms();
    CG(cugraph_graph_create_with_times_sg(h, &props, nullptr, vs, vd, vw, nullptr, nullptr, nullptr, nullptr, FALSE, TRUE, FALSE,
                                          FALSE, FALSE, FALSE, &g, &err));
    const double create_ms = now_ms() - t0;
    // sources: endpoints of edges spread over the list (never isolated)
    std::vector<int32_t> sources(n_src + 1);
    for (int k = 0; k <= n_src; ++k)
      CK(cudaMemcpy(&sources[k], src + (long long)(k + 1) * (E / (n_src + 2)), sizeof(int32_t), cudaMemcpyDeviceToHost));
    int32_t* d_seed;
    CK(cudaMalloc(&d_seed, sizeof(int32_t)));
    for (int alg = 0; alg < 2; ++alg) {
      double sum_ms = 0, min_ms = 1e30, max_ms = 0;
      long long reached = 0;
      const int n = alg == 0 ? n_src : (n_src < 4 ? n_src : 4);
      for (int k = 0; k <= n; ++k) {  // source 0 warms up
        cugraph_paths_result_t* res = nullptr;
        CK(cudaMemcpy(d_seed, &sources[k], sizeof(int32_t), cudaMemcpyHostToDevice));
        auto* vseed = view(d_seed, 1, INT32);
        t0 = now_ms();
        if (alg == 0) CG(cugraph_bfs(h, g, vseed, TRUE, (size_t)INT_MAX - 1, TRUE, FALSE, &res, &err));
        else CG(cugraph_sssp(h, g, (size_t)sources[k], 3.4e38, TRUE, FALSE, &res, &err));
        const double ms = now_ms() - t0;
        if (k == n) {  // size of the last source's component
          auto* dv      = cugraph_paths_result_get_distances(res);
          const size_t m = cugraph_type_erased_device_array_view_size(dv);
          std::vector<int32_t> hd(m);
          CK(cudaMemcpy(hd.data(), cugraph_type_erased_device_array_view_pointer(dv), m * 4, cudaMemcpyDeviceToHost));
          for (size_t i = 0; i < m; ++i) reached += alg == 0 ? (hd[i] != INT_MAX) : (hd[i] != 0x7f7fffff);
          cugraph_type_erased_device_array_view_free(dv);
        }
        cugraph_paths_result_free(res);
        cugraph_type_erased_device_array_view_free(vseed);
        if (k > 0) {
          sum_ms += ms;
          if (ms < min_ms) min_ms = ms;
          if (ms > max_ms) max_ms = ms;
        }
      }
      std::printf("{\"mode\": \"%s\", \"scale\": %d, \"create_ms\": %.1f, \"sources\": %d, \"mean_ms\": %.3f, \"min_ms\": %.3f, "
                  "\"max_ms\": %.3f, \"reached_last\": %lld}\n",
                  alg == 0 ? "bfs" : "sssp", scale, create_ms, n, sum_ms / n, min_ms, max_ms, reached);
    }
    cugraph_graph_free(g);
  }
  cugraph_free_resource_handle(h);
  return 0;
}
