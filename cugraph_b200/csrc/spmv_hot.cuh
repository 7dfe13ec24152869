// Column-blocked pull sweep for the degree>=32 rows: gathers served from SHARED MEMORY.
//
// Why: on RMAT-24 the plain sweep (spmv.cuh) is bound by the L2 -> SM path, not by HBM: every gather
// of x[src] costs a 32-byte L2 sector for 4 useful bytes (ncu: 0.86 L2 sectors per edge, lts 67 %,
// l1tex 73 %, DRAM 16 % — profiles/r01_ncu_k_spmv_hi_v1.csv).  The source space is therefore cut into
// B hot blocks of W vertices whose x-slice (192 KiB) a persistent CTA keeps in shared memory, filled by
// TMA bulk copies (cp.async.bulk + mbarrier).  Rows keep their neighbours sorted by source id, so a
// row's adjacency is already partitioned by block.  Staging (graph_build.cu) cuts every (row, block)
// segment into PIECES of <= 64 entries, stored as <= 8 LANE SLOTS of 8 entries with 16-bit local column
// ids, and orders the pieces by (block, slots per piece).  32 consecutive pieces of one class are a
// GROUP: one warp, lane = piece.  Per step a lane reads its 8 ids with one 128-bit load (the warp reads
// 512 contiguous bytes), gathers 8 values from shared memory and adds them in fp64 into a register;
// after the group's <= 8 steps every lane issues ONE fp64 atomic into acc_hi[row] — no per-entry
// predicates (padding entries read a zero), no per-step shuffles.  Only the class of full 64-entry pieces
// can hold several pieces of one row in a group; there a segmented shuffle reduction runs first.
//
// Execution: work units (sub-units of one block, about 8192 slots) are ordered by block and every CTA owns
// a contiguous, cost-balanced range of them, so a CTA refills its shared memory only a couple of times
// per sweep (one atomic cursor for ALL CTAs made every CTA walk every block: 148 x B slice fills were
// 43 % of the shared-memory wavefronts, profiles/r01_ncu_k_spmv_blocked_v4.csv).  A CTA that runs out
// of units steals from the ranges of the following CTAs (per-range atomic cursors; the next unit is
// fetched while the current one is processed).  The cold block (sources >= B*W) runs through the same
// code with global gathers.
#pragma once
#include "spmv.cuh"

namespace b200 {

constexpr int kHotThreads  = 1024;
constexpr int kHotWarps    = kHotThreads / 32;
constexpr int kHotDynSmem  = kHotSliceBytes;
constexpr int kHotTmaPiece = 16 * 1024;  // bytes per bulk copy of the slice

#ifndef B200_HOST_EMU
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "WAIT_LOOP:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra DONE;\n"
    "bra WAIT_LOOP;\n"
    "DONE:\n"
    "}\n" ::"r"(smem_u32(bar)),
    "r"(parity)
    : "memory");
}
// TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                 smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ uint4 ld_stream_v4(const void* p)
{
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
#else  // host emulation (emu/cuda_runtime.h): a bulk copy is a memcpy by the issuing thread, waiting on the mbarrier is a
       // CTA barrier (every thread of the CTA waits on it in these kernels)
inline void mbar_init(uint64_t*, unsigned) {}
inline void mbar_expect_tx(uint64_t*, unsigned) {}
inline void mbar_wait(uint64_t*, unsigned) { __syncthreads(); }
inline void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t*) { std::memcpy(dst_smem, src_gmem, bytes); }
inline uint4 ld_stream_v4(const void* p)
{
  uint4 v;
  std::memcpy(&v, p, sizeof(v));
  return v;
}
#endif

// the dynamic shared memory of a kernel
#ifndef B200_HOST_EMU
#define B200_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#else
#define B200_DYN_SMEM(name) extern unsigned char name[]  /* one CTA at a time: emu/emu_debug.cpp defines b200::smem_raw */
#endif

// consecutive groups of one class (mirrored by hot_sub_host_t, graph_build.cu)
struct hot_sub_t {
  int32_t slot_begin;  // first slot of group 0; group q starts at slot_begin + q * 32 * cls
  int32_t row_begin;   // seg_row index of (group 0, lane 0)
  int32_t n_groups;
  int32_t cls;         // slots per piece = steps per group, 1..8
};
// a work unit: consecutive sub-units of one block (mirrored by hot_unit_host_t)
struct hot_unit_t {
  int32_t sub_begin;
  int32_t sub_end;
  int32_t block;
  int32_t pad;
};

// the 8 column ids of a lane slot: one 128-bit load (hot, 16-bit ids) or two (cold, 32-bit ids)
struct slot_ids_t {
  uint4 a, b;
};
template <bool HOT>
__device__ __forceinline__ slot_ids_t hot_slot_load(int s, uint16_t const* __restrict__ idx16,
                                                    int32_t const* __restrict__ idx32, int cold_slot0)
{
  slot_ids_t r;
  if (HOT) {
    r.a = ld_stream_v4(idx16 + (size_t)(unsigned)s * kHotSlot);
    r.b = r.a;
  } else {
    r.a = ld_stream_v4(idx32 + (size_t)(unsigned)(s - cold_slot0) * kHotSlot);
    r.b = ld_stream_v4(idx32 + (size_t)(unsigned)(s - cold_slot0) * kHotSlot + 4);
  }
  return r;
}

// sum of the 8 entries of a lane slot (fp64)
template <typename T, bool WEIGHTED, bool HOT>
__device__ __forceinline__ double hot_slot_sum(slot_ids_t const& ids, int s, T const* __restrict__ w,
                                               T const* __restrict__ x, T const* __restrict__ sx)
{
  T v[kHotSlot];
  if (HOT) {
    // byte offsets into the slice: (id16 << 2), extracted with one shift + one mask each
    const unsigned m = 0x3fffcu;
    const char* base = reinterpret_cast<const char*>(sx);
    v[0] = *reinterpret_cast<const T*>(base + (((ids.a.x << 2) & m) * (sizeof(T) / 4)));
    v[1] = *reinterpret_cast<const T*>(base + (((ids.a.x >> 14) & m) * (sizeof(T) / 4)));
    v[2] = *reinterpret_cast<const T*>(base + (((ids.a.y << 2) & m) * (sizeof(T) / 4)));
    v[3] = *reinterpret_cast<const T*>(base + (((ids.a.y >> 14) & m) * (sizeof(T) / 4)));
    v[4] = *reinterpret_cast<const T*>(base + (((ids.a.z << 2) & m) * (sizeof(T) / 4)));
    v[5] = *reinterpret_cast<const T*>(base + (((ids.a.z >> 14) & m) * (sizeof(T) / 4)));
    v[6] = *reinterpret_cast<const T*>(base + (((ids.a.w << 2) & m) * (sizeof(T) / 4)));
    v[7] = *reinterpret_cast<const T*>(base + (((ids.a.w >> 14) & m) * (sizeof(T) / 4)));
  } else {
    v[0] = x[ids.a.x]; v[1] = x[ids.a.y]; v[2] = x[ids.a.z]; v[3] = x[ids.a.w];
    v[4] = x[ids.b.x]; v[5] = x[ids.b.y]; v[6] = x[ids.b.z]; v[7] = x[ids.b.w];
  }
  if (WEIGHTED) {
#pragma unroll
    for (int k = 0; k < kHotSlot; ++k) v[k] *= ld_stream(w + (size_t)(unsigned)s * kHotSlot + k);
  }
  return (((double)v[0] + (double)v[1]) + ((double)v[2] + (double)v[3])) +
         (((double)v[4] + (double)v[5]) + ((double)v[6] + (double)v[7]));
}

// end of a group: one fp64 atomic per lane.  SAME_ROW_RUNS (class of full pieces): consecutive lanes may
// hold pieces of the same row — suffix-sum inside the runs first, run heads emit.
template <bool SAME_ROW_RUNS>
__device__ __forceinline__ void hot_emit(double acc, int row, double* __restrict__ acc_hi, int lane)
{
#ifdef B200_HOST_EMU  // called lane by lane outside a launch (emu_debug.cpp: model_blocked): no warp to reduce with
  if (!emu::in_fiber()) {
    if (row >= 0) atomicAdd(acc_hi + row, acc);
    return;
  }
#endif
  if (SAME_ROW_RUNS) {
    const int r0 = __shfl_sync(0xffffffffu, row, 0);
    if (__all_sync(0xffffffffu, row == r0)) {  // 32 pieces of one hub row
      acc = warp_sum(acc);
      if (lane == 0 && r0 >= 0) atomicAdd(acc_hi + r0, acc);
      return;
    }
    const int left = __shfl_up_sync(0xffffffffu, row, 1);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double nb = __shfl_down_sync(0xffffffffu, acc, o);
      const int rn    = __shfl_down_sync(0xffffffffu, row, o);
      if (lane + o < 32 && rn == row) acc += nb;
    }
    if (lane > 0 && left == row) row = -1;  // not the head of its run
  }
  if (row >= 0) atomicAdd(acc_hi + row, acc);
}

// all groups of one sub-unit that fall to this warp (q = q0, q0 + 32, ...); the first ids / row of the
// next group are requested before the current group is reduced
template <typename T, bool WEIGHTED, bool HOT>
__device__ __forceinline__ void hot_run_groups(hot_sub_t const sb, int q, int lane, int32_t const* __restrict__ seg_row,
                                               uint16_t const* __restrict__ idx16, int32_t const* __restrict__ idx32,
                                               int cold_slot0, T const* __restrict__ w, T const* __restrict__ x,
                                               T const* __restrict__ sx, double* __restrict__ acc_hi)
{
  if (q >= sb.n_groups) return;
  const int cls    = sb.cls;
  const int gslots = 32 * cls;  // slot numbers fit 31 bits (checked at staging)
  int s            = sb.slot_begin + q * gslots + lane;
  int ri           = sb.row_begin + q * 32 + lane;
  slot_ids_t ids         = hot_slot_load<HOT>(s, idx16, idx32, cold_slot0);
  int row                = ld_stream(seg_row + ri);
  while (true) {
    double acc = 0.0;
    for (int j = 1; j < cls; ++j) {
      const slot_ids_t nx = hot_slot_load<HOT>(s + 32 * j, idx16, idx32, cold_slot0);
      acc += hot_slot_sum<T, WEIGHTED, HOT>(ids, s + 32 * (j - 1), w, x, sx);
      ids = nx;
    }
    q += 32;
    const bool more = q < sb.n_groups;
    slot_ids_t nx   = ids;
    int nrow        = -1;
    if (more) {
      nx   = hot_slot_load<HOT>(s + 32 * gslots, idx16, idx32, cold_slot0);
      nrow = ld_stream(seg_row + ri + 1024);
    }
    acc += hot_slot_sum<T, WEIGHTED, HOT>(ids, s + 32 * (cls - 1), w, x, sx);
    if (cls == kHotSlot) hot_emit<true>(acc, row, acc_hi, lane);
    else hot_emit<false>(acc, row, acc_hi, lane);
    if (!more) break;
    s += 32 * gslots;
    ri += 1024;
    ids = nx;
    row = nrow;
  }
}

#ifndef B200_HOST_EMU
__device__ __forceinline__ int ld_volatile(const int* p)
{
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
#else
inline int ld_volatile(const int* p) { return *p; }
#endif

// next unit for this CTA (called by all lanes of warp 0): own range first, then the ranges of the
// following CTAs.  victim_off = how many ranges (starting with the own one) are known to be exhausted.
__device__ __forceinline__ int hot_fetch_unit(int* __restrict__ cursor, int32_t const* __restrict__ cta_range, int n_cta,
                                              int n_units, int& victim_off, int lane)
{
  while (victim_off < n_cta) {
    int v = (int)blockIdx.x + victim_off;
    if (v >= n_cta) v -= n_cta;
    int u = -1;
    if (lane == 0) {
      u = cta_range[v] + atomicAdd(cursor + v, 1);
      if (u >= cta_range[v + 1]) u = -1;
    }
    u = __shfl_sync(0xffffffffu, u, 0);
    if (u >= 0) return u;
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
  return n_units;
}

template <typename T, bool WEIGHTED>
__global__ void __launch_bounds__(kHotThreads, 1)
k_spmv_blocked(hot_unit_t const* __restrict__ units, int n_units, int* __restrict__ unit_counter,
               int32_t const* __restrict__ cta_range, hot_sub_t const* __restrict__ subs,
               int32_t const* __restrict__ seg_row, uint16_t const* __restrict__ idx16,
               int32_t const* __restrict__ idx32, int cold_slot0, T const* __restrict__ w,
               T const* __restrict__ x, double* __restrict__ acc_hi, int W, int B, pr_state_t const* __restrict__ st)
{
  B200_DYN_SMEM(smem_raw);
  T* sx = reinterpret_cast<T*>(smem_raw);
  __shared__ uint64_t bar;
  __shared__ int s_next;
  if (st->done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int victim_off = 0;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  if (warp == 0) {
    const int n = hot_fetch_unit(unit_counter, cta_range, (int)gridDim.x, n_units, victim_off, lane);
    if (lane == 0) s_next = n;
  }
  if (threadIdx.x < kHotZeroPad) sx[W + threadIdx.x] = (T)0;  // the padding column(s) of every slice
  unsigned phase = 0;
  int cur_block  = -1;
  while (true) {
    __syncthreads();  // s_next is published; everyone is done with the previous unit's slice
    const int u = s_next;
    __syncthreads();
    if (u >= n_units) break;
    if (warp == 0) {  // fetch the next unit while working
      const int n = hot_fetch_unit(unit_counter, cta_range, (int)gridDim.x, n_units, victim_off, lane);
      if (lane == 0) s_next = n;
    }
    const hot_unit_t un = units[u];
    const int b         = un.block;
    const bool hot      = b < B;
    if (hot && b != cur_block) {
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
    // groups are dealt round-robin to the warps, continuing across the sub-units of the unit
    int dealt = 0;
    for (int si = un.sub_begin; si < un.sub_end; ++si) {
      const hot_sub_t sb = subs[si];
      const int q0       = (warp - dealt) & (kHotWarps - 1);
      dealt += sb.n_groups;
      if (hot) hot_run_groups<T, WEIGHTED, true>(sb, q0, lane, seg_row, idx16, idx32, cold_slot0, w, x, sx, acc_hi);
      else hot_run_groups<T, WEIGHTED, false>(sb, q0, lane, seg_row, idx16, idx32, cold_slot0, w, x, sx, acc_hi);
    }
  }
}

// y[row] = acc * alpha + init for every degree>=32 row; clears the accumulators and the unit counter
template <typename T>
__global__ void k_spmv_blocked_finish(double* __restrict__ acc_hi, int n_hi, T* __restrict__ y,
                                      int32_t const* __restrict__ row_vertex, double alpha, int* unit_counter,
                                      int n_cta, pr_state_t const* __restrict__ st)
{
  if (st->done) return;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_cta) unit_counter[r] = 0;
  if (r >= n_hi) return;
  y[row_vertex ? row_vertex[r] : r] = (T)(acc_hi[r] * alpha + st->init);
  acc_hi[r]                          = 0.0;
}

// EXPERIMENTAL (CUGRAPH_B200_LOW_ELL=2): the ELL sweep of the degree < 32 rows as a persistent kernel whose CTAs keep
// x[0, W) — the first column block, i.e. the sources with the largest in-degree — in shared memory (one TMA fill per
// CTA and sweep); gathers of those sources are served from shared memory, the rest from L2 as before.
template <typename T, bool WEIGHTED>
__global__ void __launch_bounds__(kHotThreads, 1)
k_spmv_low_ell_hot(int32_t const* __restrict__ ell, T const* __restrict__ ellw, T const* __restrict__ x, T* __restrict__ y,
                   int32_t const* __restrict__ row_vertex, low_ell_args_t L, int W, double alpha,
                   pr_state_t const* __restrict__ st)
{
  B200_DYN_SMEM(smem_raw);
  T* sx = reinterpret_cast<T*>(smem_raw);
  __shared__ uint64_t bar;
  if (st->done) return;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    const unsigned bytes = (unsigned)(W * sizeof(T));
    mbar_expect_tx(&bar, bytes);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(x);
    for (unsigned o = 0; o < bytes; o += kHotTmaPiece)
      tma_bulk_g2s(smem_raw + o, src + o, (bytes - o) < (unsigned)kHotTmaPiece ? (bytes - o) : (unsigned)kHotTmaPiece, &bar);
  }
  __syncthreads();  // the barrier is initialised before anyone waits on it
  mbar_wait(&bar, 0);
  gather_hot_t<T> g{x, sx, W};
  const double init  = st->init;
  const int n_vblock = L.block_begin[32];
  const int sub      = threadIdx.x >> 8;  // four virtual 256-thread blocks per CTA
  for (int vb = (int)blockIdx.x * 4 + sub; vb < n_vblock; vb += (int)gridDim.x * 4)
    low_ell_block<T, WEIGHTED>(vb, (int)(threadIdx.x & 255), ell, ellw, g, y, row_vertex, L, alpha, init);
}

template <typename T>
void launch_low_rows_ell_hot(handle_impl const& h, csx_t const& c, low_ell_t const& E, T const* x, T* y, double alpha,
                             pr_state_t const* st, int min_degree_covered = 32)
{
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(k_spmv_low_ell_hot<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHotDynSmem));
    CUDA_TRY(cudaFuncSetAttribute(k_spmv_low_ell_hot<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHotDynSmem));
    attr_set = true;
  }
  low_ell_args_t a = make_low_ell_args(E, min_degree_covered);
  const int blocks = a.block_begin[32];
  if (blocks <= 0) return;
  const int W    = (int)(kHotSliceBytes / sizeof(T)) - kHotZeroPad;  // x holds padded_x_elems(): reading W entries is safe
  const int grid = std::min(h.sm_count, (blocks + 3) / 4);
  if (E.w.data())
    B200_LAUNCH(h, (k_spmv_low_ell_hot<T, true>), grid, kHotThreads, kHotDynSmem, E.idx.as<int32_t>(), E.w.as<T>(), x, y,
                c.row_vertex.as<int32_t>(), a, W, alpha, st);
  else
    B200_LAUNCH(h, (k_spmv_low_ell_hot<T, false>), grid, kHotThreads, kHotDynSmem, E.idx.as<int32_t>(), E.w.as<T>(), x, y,
                c.row_vertex.as<int32_t>(), a, W, alpha, st);
}

// EXPERIMENTAL (CUGRAPH_B200_LOW_ASYNC=1): run the degree < 32 rows on the handle's second stream, concurrently with
// the persistent kernel (disjoint rows of y; both only read x and the loop state)
inline bool low_async()
{
  const char* e = std::getenv("CUGRAPH_B200_LOW_ASYNC");
  return e && std::atoi(e) != 0;
}

template <typename O, typename T>
void launch_low_rows(handle_impl const& h, csx_t const& c, T const* x, T* y, double alpha, pr_state_t const* st,
                     int first_bin = 0)
{
  if (low_ell_t const* E = low_ell_layout(h, c, sizeof(T))) {  // experimental, CUGRAPH_B200_LOW_ELL=1 | 2
    const int covered = kSegThreshold[first_bin];  // rows of degree >= this belong to the piece layout
    if (low_ell_mode() >= 2) launch_low_rows_ell_hot<T>(h, c, *E, x, y, alpha, st, covered);
    else launch_low_rows_ell<T>(h, c, *E, x, y, alpha, st, covered);
    return;
  }
  low_bins_t bins = make_low_bins(c, first_bin);
  int lblocks     = bins.block_begin[kNumSeg - 1];
  if (lblocks <= 0) return;
  if (c.weights.data())
    B200_LAUNCH(h, (k_spmv_low<O, T, true>), lblocks, 256, 0, c.offsets.as<O>(), c.indices.as<int32_t>(),
                c.weights.as<T>(), x, y, c.row_vertex.as<int32_t>(), bins, alpha, st, low_mode());
  else
    B200_LAUNCH(h, (k_spmv_low<O, T, false>), lblocks, 256, 0, c.offsets.as<O>(), c.indices.as<int32_t>(),
                c.weights.as<T>(), x, y, c.row_vertex.as<int32_t>(), bins, alpha, st, low_mode());
}

// x must hold padded_x_elems() elements, zero behind n_vertices (slices are copied whole; the cold
// block's padding entries read x[n_vertices])
template <typename O, typename T>
void launch_pull_sweep_blocked(handle_impl const& h, csx_t const& c, hot_layout_t const& L, T const* x, T* y,
                               double* acc_hi, double alpha, pr_state_t const* st)
{
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(k_spmv_blocked<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHotDynSmem));
    CUDA_TRY(cudaFuncSetAttribute(k_spmv_blocked<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHotDynSmem));
    attr_set = true;
  }
  const int grid = L.n_cta;
  const bool la  = low_async() && h.aux_stream != nullptr;  // experimental: low rows on the second stream
  if (la) CUDA_TRY(cudaEventRecord(h.ev_a, h.stream));      // x and the loop state are ready here
  if (L.slot_w.data())
    B200_LAUNCH(h, (k_spmv_blocked<T, true>), grid, kHotThreads, kHotDynSmem, L.units.as<hot_unit_t>(), L.n_units,
                L.unit_counter.as<int>(), L.cta_range.as<int32_t>(), L.subs.as<hot_sub_t>(), L.seg_row.as<int32_t>(), L.slot_idx16.as<uint16_t>(), L.slot_idx32.as<int32_t>(),
                (int)L.n_hot_slots, L.slot_w.as<T>(), x, acc_hi, L.W, L.B, st);
  else
    B200_LAUNCH(h, (k_spmv_blocked<T, false>), grid, kHotThreads, kHotDynSmem, L.units.as<hot_unit_t>(), L.n_units,
                L.unit_counter.as<int>(), L.cta_range.as<int32_t>(), L.subs.as<hot_sub_t>(), L.seg_row.as<int32_t>(), L.slot_idx16.as<uint16_t>(), L.slot_idx32.as<int32_t>(),
                (int)L.n_hot_slots, L.slot_w.as<T>(), x, acc_hi, L.W, L.B, st);
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

// (the dispatcher launch_pull_sweep_auto lives in spmv_hot_x.cuh, next to the experimental kernel variant)

// elements an x buffer needs: whole slices are TMA-copied and x[n_vertices] must be a readable zero.
// The buffer must be zero-filled once at allocation; only [0, n_vertices) is ever written afterwards.
inline size_t padded_x_elems(int32_t n_vertices, size_t elem_size)
{
  const size_t slice = kHotSliceBytes / elem_size;
  const size_t W     = slice - kHotZeroPad;
  return ((size_t)n_vertices / W + 2) * slice;
}

}  // namespace b200
