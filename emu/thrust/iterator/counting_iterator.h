// host emulation (see emu/cuda_runtime.h)
#pragma once
namespace thrust {
template <typename T>
struct counting_iterator {
  T base;
  explicit counting_iterator(T b = T()) : base(b) {}
  T operator[](long long i) const { return (T)(base + i); }
};
}  // namespace thrust
