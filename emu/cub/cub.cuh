// host emulation of the CUB device-wide primitives used by graph staging (see emu/cuda_runtime.h)
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

namespace cub {

namespace detail {
template <typename K>
inline unsigned long long digits(K k, int b0, int b1)
{
  unsigned long long v = 0;
  static_assert(sizeof(K) <= 8, "key width");
  std::memcpy(&v, &k, sizeof(K));
  const int w = b1 - b0;
  v >>= b0;
  return w >= 64 ? v : (v & ((1ull << w) - 1));
}
}  // namespace detail

struct DeviceRadixSort {
  template <typename K, typename N>
  static cudaError_t SortKeys(void* tmp, size_t& bytes, const K* in, K* out, N n, int b0, int b1, cudaStream_t = nullptr)
  {
    if (!tmp) { bytes = 1; return cudaSuccess; }
    std::vector<K> v(in, in + n);
    std::stable_sort(v.begin(), v.end(), [&](K a, K b) { return detail::digits(a, b0, b1) < detail::digits(b, b0, b1); });
    std::copy(v.begin(), v.end(), out);
    return cudaSuccess;
  }
  template <typename K, typename V, typename N>
  static cudaError_t SortPairs(void* tmp, size_t& bytes, const K* kin, K* kout, const V* vin, V* vout, N n, int b0, int b1,
                               cudaStream_t = nullptr)
  {
    if (!tmp) { bytes = 1; return cudaSuccess; }
    std::vector<long long> p((size_t)n);
    std::iota(p.begin(), p.end(), 0ll);
    std::stable_sort(p.begin(), p.end(),
                     [&](long long a, long long b) { return detail::digits(kin[a], b0, b1) < detail::digits(kin[b], b0, b1); });
    std::vector<K> ks((size_t)n);
    std::vector<V> vs((size_t)n);
    for (size_t i = 0; i < (size_t)n; ++i) { ks[i] = kin[p[i]]; vs[i] = vin[p[i]]; }
    std::copy(ks.begin(), ks.end(), kout);
    std::copy(vs.begin(), vs.end(), vout);
    return cudaSuccess;
  }
};

struct DeviceScan {
  template <typename In, typename Out, typename N>
  static cudaError_t ExclusiveSum(void* tmp, size_t& bytes, In in, Out out, N n, cudaStream_t = nullptr)
  {
    if (!tmp) { bytes = 1; return cudaSuccess; }
    auto run = in[0];
    run      = 0;
    for (N i = 0; i < n; ++i) { auto v = in[i]; out[i] = run; run += v; }
    return cudaSuccess;
  }
};

struct DeviceSelect {
  template <typename In, typename Flag, typename Out, typename Cnt, typename N>
  static cudaError_t Flagged(void* tmp, size_t& bytes, In in, Flag flags, Out out, Cnt count, N n, cudaStream_t = nullptr)
  {
    if (!tmp) { bytes = 1; return cudaSuccess; }
    long long m = 0;
    for (N i = 0; i < n; ++i)
      if (flags[i]) out[m++] = in[i];
    *count = m;
    return cudaSuccess;
  }
  template <typename In, typename Out, typename Cnt, typename N>
  static cudaError_t Unique(void* tmp, size_t& bytes, In in, Out out, Cnt count, N n, cudaStream_t = nullptr)
  {
    if (!tmp) { bytes = 1; return cudaSuccess; }
    long long m = 0;
    for (N i = 0; i < n; ++i)
      if (i == 0 || !(in[i] == in[i - 1])) out[m++] = in[i];
    *count = m;
    return cudaSuccess;
  }
};

}  // namespace cub
