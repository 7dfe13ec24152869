/*
 * bench_ref.c — the CPU arm of bench.py (`--impl reference`, `cpu_baseline`): the reference's PageRank ALGORITHM
 * (pagerank_reference, cpp/tests/link_analysis/pagerank_test.cpp:33-121; driver cpp/src/link_analysis/pagerank_impl.cuh:224-327)
 * restated for float32 storage and all host cores, plus a host RMAT generator and a parallel COO -> CSC so that the
 * BENCHMARK CONFIGURATION ITSELF (RMAT scale-24 edge-factor-16, 100 iterations) is what gets timed.
 * TEST / BENCH INFRASTRUCTURE ONLY, like everything under oracle/: the product never links it.
 *
 * The arithmetic per iteration is the one the GPU path does: x = pr / out_degree, dangling sum,
 * y[v] = (dangling * alpha + 1 - alpha) / V + alpha * sum_{u->v} x[u], row sums accumulated in double, values stored as
 * float.  libcugraph itself cannot be built here, so this port is the reference arm (`kind: "port"`).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline uint64_t mix64(uint64_t z)
{
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

/* RMAT edges with the sampling rule of the reference generator (cpp/src/generators/generate_rmat_edgelist.cuh:66-108):
 * per bit two uniforms r0, r1; src_bit = r0 > a+b; dst_bit = r1 > (src_bit ? c/(1-(a+b)) : a/(a+b)).  Counter-based
 * uniforms (one 64-bit hash per edge and bit), so the edge list is a function of (seed, edge index) whatever the thread
 * count.  Vertex ids are scrambled by a fixed bijection of [0, 2^scale) (an odd multiplier and an xor-shift), the role of
 * scramble.cuh:44-67: hubs are spread over the id range. */
void bench_rmat_edges(int scale, int64_t num_edges, uint64_t seed, double a, double b, double c, int32_t* src, int32_t* dst)
{
  const float ab = (float)(a + b), a_norm = (float)(a / (a + b)), c_norm = (float)(c / (1.0 - (a + b)));
  const uint32_t mask = (scale >= 32) ? 0xffffffffu : ((1u << scale) - 1u);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < num_edges; ++e) {
    uint32_t s = 0, d = 0;
    for (int bit = scale - 1; bit >= 0; --bit) {
      const uint64_t r = mix64(seed ^ ((uint64_t)e * 64ull + (uint64_t)bit));
      const float r0 = (float)(r >> 40) * (1.0f / 16777216.0f), r1 = (float)((r >> 8) & 0xffffffu) * (1.0f / 16777216.0f);
      const uint32_t sb = r0 > ab;
      const uint32_t db = r1 > (sb ? c_norm : a_norm);
      s |= sb << bit;
      d |= db << bit;
    }
    s = (s * 0x9e3779b1u) & mask; s ^= s >> (scale / 2 + 1); s = (s * 0x85ebca6bu) & mask;
    d = (d * 0x9e3779b1u) & mask; d ^= d >> (scale / 2 + 1); d = (d * 0x85ebca6bu) & mask;
    src[e] = (int32_t)s;
    dst[e] = (int32_t)d;
  }
}

/* COO -> CSC (offsets over destinations, indices = sources), parallel counting sort; the order inside a row is whatever the
 * threads produce (irrelevant to a sum).  out_degree[u] = number of edges leaving u. */
int bench_build_csc(int64_t num_edges, int32_t num_vertices, const int32_t* src, const int32_t* dst, int64_t* offsets /* V+1 */,
                    int32_t* indices /* E */, int32_t* out_degree /* V */)
{
  memset(offsets, 0, sizeof(int64_t) * ((size_t)num_vertices + 1));
  memset(out_degree, 0, sizeof(int32_t) * (size_t)num_vertices);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < num_edges; ++e) {
#pragma omp atomic
    offsets[dst[e] + 1]++;
#pragma omp atomic
    out_degree[src[e]]++;
  }
  for (int32_t v = 0; v < num_vertices; ++v) offsets[v + 1] += offsets[v];
  int64_t* cursor = (int64_t*)malloc(sizeof(int64_t) * (size_t)num_vertices);
  if (!cursor) return -2;
  memcpy(cursor, offsets, sizeof(int64_t) * (size_t)num_vertices);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < num_edges; ++e) {
    const int64_t p = __atomic_fetch_add(&cursor[dst[e]], 1, __ATOMIC_RELAXED);
    indices[p]      = src[e];
  }
  free(cursor);
  return 0;
}

/* `iterations` PageRank power iterations in float32 storage, starting from pr (pass 1/V), epsilon = 0 (never converges:
 * the benchmark protocol).  x and y are caller-provided scratch of V floats. */
void bench_pagerank_f32(const int64_t* offsets, const int32_t* indices, const int32_t* out_degree, int32_t num_vertices,
                        double alpha, int iterations, float* pr, float* x, float* y)
{
  const int32_t V = num_vertices;
  for (int it = 0; it < iterations; ++it) {
    double dangling = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : dangling)
    for (int32_t v = 0; v < V; ++v) {
      if (out_degree[v] == 0) {
        dangling += (double)pr[v];
        x[v] = pr[v];
      } else {
        x[v] = pr[v] / (float)out_degree[v];
      }
    }
    const double init = (dangling * alpha + (1.0 - alpha)) / (double)V;
#pragma omp parallel for schedule(dynamic, 2048)
    for (int32_t v = 0; v < V; ++v) {
      double acc = 0.0;
      for (int64_t j = offsets[v]; j < offsets[v + 1]; ++j) acc += (double)x[indices[j]];
      y[v] = (float)(acc * alpha + init);
    }
    float* t = pr;  /* the caller's pr holds the result after an even number of swaps: copy back below */
    memcpy(t, y, sizeof(float) * (size_t)V);
  }
}

int bench_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
