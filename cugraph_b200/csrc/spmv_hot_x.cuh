// EXPERIMENTAL variant of the blocked pull-sweep kernel (spmv_hot.cuh), selected with CUGRAPH_B200_HOT_X=1.
// Kept as a SEPARATE kernel: compiled into k_spmv_blocked as run-time options, the two additions below cost
// the default path 14 % (register spills under the 64-register cap), profiles/r01_notes.md.
//   * CUGRAPH_B200_HOT_C1 (default 1 here): the class of one-slot pieces keeps four groups in flight per warp
//   * CUGRAPH_B200_HOT_CLAIM (default 4 here): a CTA claims that many units of its own range per atomic and
//     runs them without a CTA barrier in between
//   * CUGRAPH_B200_HOT_NARROW=1 (layout option, graph_build.cu; unweighted graphs): pieces of <= 4 / <= 2 entries are
//     stored in 8- / 4- / 2-byte slots (classes 16 / 32 / 64: 3-4, 2, 1 entries) instead of a padded 16-byte slot; only this
//     kernel reads them
// Measured inside the combined kernel (RMAT-24, same binary): baseline 0.522 ms, C1 0.503, CLAIM=4 0.490,
// both 0.465.  To be measured as a kernel of its own next.
#pragma once
#include "spmv_hot.cuh"

namespace b200 {

// class of one-slot pieces (the bulk of the pieces once every column block is hot): one step per group, so
// there is nothing to pipeline inside a group.  The warp keeps FOUR groups in flight instead: 4 id vectors + 4
// rows are requested before the first gather (ncu on the one-group-ahead version: 22 % of all stall samples
// sat on the move that consumes the prefetched ids, profiles/r01_ncu_k_spmv_blocked_v5.csv).
template <typename T, bool WEIGHTED, bool HOT>
__device__ __forceinline__ void hot_run_groups_c1(hot_sub_t const sb, int q, int lane, int32_t const* __restrict__ seg_row,
                                                  uint16_t const* __restrict__ idx16, int32_t const* __restrict__ idx32,
                                                  int cold_slot0, T const* __restrict__ w, T const* __restrict__ x,
                                                  T const* __restrict__ sx, double* __restrict__ acc_hi)
{
  constexpr int K = HOT ? 4 : 2;  // cold ids are 32-bit: twice the registers per slot
  for (; q < sb.n_groups; q += 32 * K) {
    slot_ids_t ids[K];
    int row[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int qk = q + 32 * k;
      row[k]       = -1;
      if (qk < sb.n_groups) {  // warp-uniform
        ids[k] = hot_slot_load<HOT>(sb.slot_begin + qk * 32 + lane, idx16, idx32, cold_slot0);
        row[k] = ld_stream(seg_row + sb.row_begin + qk * 32 + lane);
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int qk = q + 32 * k;
      if (qk < sb.n_groups) {
        const double acc = hot_slot_sum<T, WEIGHTED, HOT>(ids[k], sb.slot_begin + qk * 32 + lane, w, x, sx);
        if (row[k] >= 0) atomicAdd(acc_hi + row[k], acc);
      }
    }
  }
}

#ifndef B200_HOST_EMU
__device__ __forceinline__ unsigned ld_stream_u32(const uint32_t* p)
{
  unsigned v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned ld_stream_u16(const uint16_t* p)
{
  unsigned short v;
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ld_stream_v2(const uint2* p)
{
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
#else
inline unsigned ld_stream_u32(const uint32_t* p) { return *p; }
inline unsigned ld_stream_u16(const uint16_t* p) { return *p; }
inline uint2 ld_stream_v2(const uint2* p) { return *p; }
#endif

// value of the slice entry addressed by the low (HI = false) or high half of a packed pair of 16-bit ids
template <typename T, bool HI>
__device__ __forceinline__ T hot_gather16(unsigned pair, T const* __restrict__ sx)
{
  const unsigned off = (HI ? (pair >> 14) : (pair << 2)) & 0x3fffcu;
  return *reinterpret_cast<const T*>(reinterpret_cast<const char*>(sx) + off * (sizeof(T) / 4));
}

// narrow classes (hot blocks, unweighted): one step per group, WIDTH = 1, 2 or 4 ids per lane slot; 8 / 8 / 4 groups in flight
template <typename T, int WIDTH>
__device__ __forceinline__ void hot_run_groups_narrow(hot_sub_t const sb, int q, int lane, int32_t const* __restrict__ seg_row,
                                                      uint2 const* __restrict__ idx_h, uint32_t const* __restrict__ idx_q,
                                                      uint16_t const* __restrict__ idx_s, T const* __restrict__ sx,
                                                      double* __restrict__ acc_hi)
{
  constexpr int K = WIDTH <= 2 ? 8 : 4;  // groups in flight: a quarter / single slot costs one id register per group
  for (; q < sb.n_groups; q += 32 * K) {
    uint2 ids[K];
    int row[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int qk = q + 32 * k;
      row[k]       = -1;
      ids[k]       = make_uint2(0u, 0u);
      if (qk < sb.n_groups) {  // warp-uniform
        const size_t s = (size_t)(unsigned)(sb.slot_begin + qk * 32 + lane);
        if (WIDTH == 4) ids[k] = ld_stream_v2(idx_h + s);
        else if (WIDTH == 2) ids[k].x = ld_stream_u32(idx_q + s);
        else ids[k].x = ld_stream_u16(idx_s + s);
        row[k] = ld_stream(seg_row + sb.row_begin + qk * 32 + lane);
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (q + 32 * k < sb.n_groups) {
        double acc = (double)hot_gather16<T, false>(ids[k].x, sx);
        if (WIDTH >= 2) acc += (double)hot_gather16<T, true>(ids[k].x, sx);
        if (WIDTH == 4) acc += (double)hot_gather16<T, false>(ids[k].y, sx) + (double)hot_gather16<T, true>(ids[k].y, sx);
        if (row[k] >= 0) atomicAdd(acc_hi + row[k], acc);
      }
    }
  }
}

// next units for this CTA (called by all lanes of warp 0): own range first — `claim` consecutive units per
// atomic, they are processed without a CTA barrier in between — then single units of the following CTAs'
// ranges.  victim_off = how many ranges (starting with the own one) are known to be exhausted.
// Returns the first unit (n_units when everything is done) and sets count.
__device__ __forceinline__ int hot_fetch_units(int* __restrict__ cursor, int32_t const* __restrict__ cta_range, int n_cta,
                                              int n_units, int claim, int& victim_off, int& count, int lane)
{
  while (victim_off < n_cta) {
    int v = (int)blockIdx.x + victim_off;
    if (v >= n_cta) v -= n_cta;
    const int want = victim_off == 0 ? claim : 1;
    int u = -1, c = 0;
    if (lane == 0) {
      u             = cta_range[v] + atomicAdd(cursor + v, want);
      const int end = cta_range[v + 1];
      c             = end - u < want ? end - u : want;
      if (c <= 0) u = -1;
    }
    u = __shfl_sync(0xffffffffu, u, 0);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (u >= 0) {
      count = c;
      return u;
    }
    ++victim_off;
    while (victim_off < n_cta) {  // look 32 ranges ahead at a time for one that still has units
      int vv         = (int)blockIdx.x + victim_off + lane;
      const bool inr = victim_off + lane < n_cta;
      if (vv >= n_cta) vv -= n_cta;
      const bool has   = inr && ld_volatile(cursor + vv) < cta_range[vv + 1] - cta_range[vv];
      const unsigned m = __ballot_sync(0xffffffffu, has);
      if (m) {
        victim_off += __ffs(m) - 1;
        break;
      }
      victim_off += 32;
    }
  }
  count = 0;
  return n_units;
}

template <typename T, bool WEIGHTED>
__global__ void __launch_bounds__(kHotThreads, 1)
k_spmv_blocked_x(hot_unit_t const* __restrict__ units, int n_units, int* __restrict__ unit_counter,
               int32_t const* __restrict__ cta_range, hot_sub_t const* __restrict__ subs,
               int32_t const* __restrict__ seg_row, uint16_t const* __restrict__ idx16,
               int32_t const* __restrict__ idx32, int cold_slot0, T const* __restrict__ w,
               T const* __restrict__ x, double* __restrict__ acc_hi, int W, int B, int c1_wide,
               int claim, uint2 const* __restrict__ idx_h, uint32_t const* __restrict__ idx_q,
               uint16_t const* __restrict__ idx_s, pr_state_t const* __restrict__ st)
{
  B200_DYN_SMEM(smem_raw);
  T* sx = reinterpret_cast<T*>(smem_raw);
  __shared__ uint64_t bar;
  __shared__ int s_next, s_count;
  if (st->done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int victim_off = 0;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  if (warp == 0) {
    int cnt     = 0;
    const int n = hot_fetch_units(unit_counter, cta_range, (int)gridDim.x, n_units, claim, victim_off, cnt, lane);
    if (lane == 0) {
      s_next  = n;
      s_count = cnt;
    }
  }
  if (threadIdx.x < kHotZeroPad) sx[W + threadIdx.x] = (T)0;  // the padding column(s) of every slice
  unsigned phase = 0;
  int cur_block  = -1;
  while (true) {
    __syncthreads();  // s_next is published; everyone is done with the previous units' slice
    const int u0 = s_next, ucnt = s_count;
    __syncthreads();
    if (u0 >= n_units) break;
    if (warp == 0) {  // fetch the next claim while working
      int cnt     = 0;
      const int n = hot_fetch_units(unit_counter, cta_range, (int)gridDim.x, n_units, claim, victim_off, cnt, lane);
      if (lane == 0) {
        s_next  = n;
        s_count = cnt;
      }
    }
    int dealt = 0;  // groups are dealt round-robin to the warps, continuing across sub-units and units
    for (int u = u0; u < u0 + ucnt; ++u) {
      const hot_unit_t un = units[u];
      const int b         = un.block;
      const bool hot      = b < B;
      if (hot && b != cur_block) {
        if (u != u0) __syncthreads();  // warps of this claim may still gather from the old slice
        if (threadIdx.x == 0) {
          const unsigned bytes = (unsigned)(W * sizeof(T));
          mbar_expect_tx(&bar, bytes);
          const unsigned char* src = reinterpret_cast<const unsigned char*>(x + (size_t)b * W);
          for (unsigned o = 0; o < bytes; o += kHotTmaPiece)
            tma_bulk_g2s(smem_raw + o, src + o, (bytes - o) < (unsigned)kHotTmaPiece ? (bytes - o) : (unsigned)kHotTmaPiece, &bar);
        }
        cur_block = b;
        mbar_wait(&bar, phase);
        phase ^= 1;
      }
      for (int si = un.sub_begin; si < un.sub_end; ++si) {
        const hot_sub_t sb = subs[si];
        const int q0       = (warp - dealt) & (kHotWarps - 1);
        dealt += sb.n_groups;
        if (sb.cls > 8) {  // narrow classes exist only in hot blocks of narrow layouts
          if (sb.cls == 16) hot_run_groups_narrow<T, 4>(sb, q0, lane, seg_row, idx_h, idx_q, idx_s, sx, acc_hi);
          else if (sb.cls == 32) hot_run_groups_narrow<T, 2>(sb, q0, lane, seg_row, idx_h, idx_q, idx_s, sx, acc_hi);
          else hot_run_groups_narrow<T, 1>(sb, q0, lane, seg_row, idx_h, idx_q, idx_s, sx, acc_hi);
        } else if (c1_wide && sb.cls == 1 && hot) {  // the cold block (if any) stays on the generic loop
          hot_run_groups_c1<T, WEIGHTED, true>(sb, q0, lane, seg_row, idx16, idx32, cold_slot0, w, x, sx, acc_hi);
        } else if (hot) {
          hot_run_groups<T, WEIGHTED, true>(sb, q0, lane, seg_row, idx16, idx32, cold_slot0, w, x, sx, acc_hi);
        } else {
          hot_run_groups<T, WEIGHTED, false>(sb, q0, lane, seg_row, idx16, idx32, cold_slot0, w, x, sx, acc_hi);
        }
      }
    }
  }
}

inline int hot_x_enabled()
{
  const char* e = std::getenv("CUGRAPH_B200_HOT_X");
  return e ? std::atoi(e) : 0;
}
inline int hot_c1_wide()
{
  const char* e = std::getenv("CUGRAPH_B200_HOT_C1");
  return e ? std::atoi(e) : 1;
}
inline int hot_claim()
{
  const char* e = std::getenv("CUGRAPH_B200_HOT_CLAIM");
  const int c   = e ? std::atoi(e) : 4;
  return c < 1 ? 1 : (c > 64 ? 64 : c);
}

template <typename O, typename T>
void launch_pull_sweep_blocked_x(handle_impl const& h, csx_t const& c, hot_layout_t const& L, T const* x, T* y,
                                 double* acc_hi, double alpha, pr_state_t const* st)
{
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(k_spmv_blocked_x<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHotDynSmem));
    CUDA_TRY(cudaFuncSetAttribute(k_spmv_blocked_x<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHotDynSmem));
    attr_set = true;
  }
  const int grid = L.n_cta;
  const bool la  = low_async() && h.aux_stream != nullptr;  // experimental: low rows on the second stream
  if (la) CUDA_TRY(cudaEventRecord(h.ev_a, h.stream));      // x and the loop state are ready here
  if (L.slot_w.data())
    B200_LAUNCH(h, (k_spmv_blocked_x<T, true>), grid, kHotThreads, kHotDynSmem, L.units.as<hot_unit_t>(), L.n_units,
                L.unit_counter.as<int>(), L.cta_range.as<int32_t>(), L.subs.as<hot_sub_t>(), L.seg_row.as<int32_t>(),
                L.slot_idx16.as<uint16_t>(), L.slot_idx32.as<int32_t>(), (int)L.n_hot_slots, L.slot_w.as<T>(), x, acc_hi, L.W,
                L.B, hot_c1_wide(), hot_claim(), L.slot_idx_h.as<uint2>(), L.slot_idx_q.as<uint32_t>(), L.slot_idx_s.as<uint16_t>(), st);
  else
    B200_LAUNCH(h, (k_spmv_blocked_x<T, false>), grid, kHotThreads, kHotDynSmem, L.units.as<hot_unit_t>(), L.n_units,
                L.unit_counter.as<int>(), L.cta_range.as<int32_t>(), L.subs.as<hot_sub_t>(), L.seg_row.as<int32_t>(),
                L.slot_idx16.as<uint16_t>(), L.slot_idx32.as<int32_t>(), (int)L.n_hot_slots, L.slot_w.as<T>(), x, acc_hi, L.W,
                L.B, hot_c1_wide(), hot_claim(), L.slot_idx_h.as<uint2>(), L.slot_idx_q.as<uint32_t>(), L.slot_idx_s.as<uint16_t>(), st);
  if (la) {  // queued behind the persistent kernel: its blocks fill the SMs that the blocked kernel's tail frees
    CUDA_TRY(cudaStreamWaitEvent(h.aux_stream, h.ev_a, 0));
    handle_impl ha = h;
    ha.stream      = h.aux_stream;
    ha.launches    = 0;
    launch_low_rows<O, T>(ha, c, x, y, alpha, st, L.seg_k);
    h.launches += ha.launches;
    CUDA_TRY(cudaEventRecord(h.ev_b, h.aux_stream));
  }
  B200_LAUNCH(h, (k_spmv_blocked_finish<T>), (L.n_hi + 255) / 256, 256, 0, acc_hi, L.n_hi, y, c.row_vertex.as<int32_t>(),
              alpha, L.unit_counter.as<int>(), L.n_cta, st);
  if (la) CUDA_TRY(cudaStreamWaitEvent(h.stream, h.ev_b, 0));
  else launch_low_rows<O, T>(h, c, x, y, alpha, st, L.seg_k);
}

// dispatch: blocked layout when it exists for this graph, else the plain edge-balanced sweep
template <typename O, typename T>
void launch_pull_sweep_auto(handle_impl const& h, csx_t const& c, int32_t n_vertices, T const* x, T* y, double* acc_hi,
                            double alpha, pr_state_t const* st)
{
  hot_layout_t const* L = hot_layout(h, c, n_vertices, sizeof(T));
  if (!L) launch_pull_sweep<O, T>(h, c, x, y, acc_hi, alpha, st);
  else if (L->narrow || hot_x_enabled()) launch_pull_sweep_blocked_x<O, T>(h, c, *L, x, y, acc_hi, alpha, st);
  else launch_pull_sweep_blocked<O, T>(h, c, *L, x, y, acc_hi, alpha, st);
}

}  // namespace b200
