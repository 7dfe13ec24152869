// Column-blocked pull sweep for the degree>=32 rows: gathers served from SHARED MEMORY.
//
// Why: on RMAT-24 the plain sweep (spmv.cuh) is bound by the L2 -> SM path, not by HBM: every gather
// of x[src] costs a 32-byte L2 sector for 4 useful bytes (ncu: 0.86 L2 sectors per edge, lts 67 %,
// l1tex 73 %, DRAM 16 % — profiles/r01_ncu_k_spmv_hi_v1.csv).  The source space is therefore cut into
// B hot blocks of W vertices whose x-slice (192 KiB) a persistent CTA keeps in shared memory, filled by
// TMA bulk copies (cp.async.bulk + mbarrier).  Rows keep their neighbours sorted by source id, so a
// row's adjacency is already partitioned by block; staging stores the (row, block) segments block-major,
// cut into LANE SLOTS of 8 entries with 16-bit local column ids (hot_layout_t, graph.cuh): a lane reads
// its 8 ids with one 128-bit load, gathers 8 values from shared memory and adds them in fp64 — no
// per-entry predicates (padding entries read a zero), no segment walk.
//
// Execution: work units (<= 8192 slots of one block) are handed out dynamically through one atomic
// counter (the next unit is fetched while the current one is processed); a CTA refills its shared
// memory only when its next unit belongs to another block.  A warp handles 32 consecutive slots per
// step.  If the 32 slots belong to one row (hub segments) the partials are folded with shuffles into ONE
// fp64 atomic; otherwise lanes first combine with their right neighbours of the same row (3 shuffle
// steps) and the surviving heads issue one fp64 atomic each into acc_hi[row] — the only atomics on
// the path.  The cold block (sources >= B*W) runs through the same code with global gathers.
#pragma once
#include "spmv.cuh"

namespace b200 {

constexpr int kHotThreads  = 1024;
constexpr int kHotWarps    = kHotThreads / 32;
constexpr int kHotDynSmem  = kHotSliceBytes;
constexpr int kHotTmaPiece = 16 * 1024;  // bytes per bulk copy of the slice

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

// a work unit: consecutive lane slots of one block (mirrored by hot_unit_host_t, graph_build.cu)
struct hot_unit_t {
  int32_t slot_begin;
  int32_t slot_end;
  int32_t block;
  int32_t combine;  // 1: segments of this block span several slots, combine neighbours before the atomics
};

// the 8 column ids of lane slot s: one 128-bit load (hot, 16-bit ids) or two (cold, 32-bit ids)
struct slot_ids_t {
  uint4 a, b;
};
template <bool HOT>
__device__ __forceinline__ slot_ids_t hot_slot_load(long long s, bool valid, uint16_t const* __restrict__ idx16,
                                                    int32_t const* __restrict__ idx32, long long cold_slot0)
{
  slot_ids_t r;
  r.a = make_uint4(0, 0, 0, 0);
  r.b = make_uint4(0, 0, 0, 0);
  if (valid) {
    if (HOT) {
      r.a = ld_stream_v4(idx16 + s * kHotSlot);
    } else {
      r.a = ld_stream_v4(idx32 + (s - cold_slot0) * kHotSlot);
      r.b = ld_stream_v4(idx32 + (s - cold_slot0) * kHotSlot + 4);
    }
  }
  return r;
}

// sum of the 8 entries of a lane slot (fp64); an invalid slot has all-zero ids and row -1: its value is
// never emitted
template <typename T, bool WEIGHTED, bool HOT>
__device__ __forceinline__ double hot_slot_sum(slot_ids_t const& ids, long long s, bool valid, T const* __restrict__ w,
                                               T const* __restrict__ x, T const* __restrict__ sx)
{
  unsigned c[kHotSlot];
  if (HOT) {
    c[0] = ids.a.x & 0xffffu; c[1] = ids.a.x >> 16; c[2] = ids.a.y & 0xffffu; c[3] = ids.a.y >> 16;
    c[4] = ids.a.z & 0xffffu; c[5] = ids.a.z >> 16; c[6] = ids.a.w & 0xffffu; c[7] = ids.a.w >> 16;
  } else {
    c[0] = ids.a.x; c[1] = ids.a.y; c[2] = ids.a.z; c[3] = ids.a.w;
    c[4] = ids.b.x; c[5] = ids.b.y; c[6] = ids.b.z; c[7] = ids.b.w;
  }
  T v[kHotSlot];
#pragma unroll
  for (int k = 0; k < kHotSlot; ++k) v[k] = HOT ? sx[c[k]] : x[c[k]];
  if (WEIGHTED) {
#pragma unroll
    for (int k = 0; k < kHotSlot; ++k) v[k] *= valid ? ld_stream(w + s * kHotSlot + k) : (T)0;
  }
  return (((double)v[0] + (double)v[1]) + ((double)v[2] + (double)v[3])) +
         (((double)v[4] + (double)v[5]) + ((double)v[6] + (double)v[7]));
}

// fold the 32 per-slot partials of a warp step into acc_hi[row] (row < 0: nothing to emit).
// The kernel is bound by the MIO pipe (shared-memory gathers + shuffles, profiles/r01_ncu_k_spmv_blocked_v4.csv),
// so the combine uses one packed shuffle for (row, alive) and a vote instead of a shuffle for the retire
// flag, and is skipped for units whose segments are mostly single slots (COMBINE = false).
template <bool COMBINE>
__device__ __forceinline__ void hot_emit(double acc, int row, double* __restrict__ acc_hi, int lane)
{
  const int r0 = __shfl_sync(0xffffffffu, row, 0);
  if (__all_sync(0xffffffffu, row == r0)) {  // one row (hub segment, or 32 padding slots)
    acc = warp_sum(acc);
    if (lane == 0 && r0 >= 0) atomicAdd(acc_hi + r0, acc);
    return;
  }
  if (COMBINE) {
    // slots of a row are consecutive lanes: aligned runs of up to 8 collapse into their head lane
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) {
      const double nb = __shfl_down_sync(0xffffffffu, acc, o);
      const int rn    = __shfl_down_sync(0xffffffffu, row, o);  // -1 when the neighbour is retired / padding
      const bool take = ((lane & (2 * o - 1)) == 0) && row >= 0 && rn == row;
      if (take) acc += nb;
      const unsigned takes = __ballot_sync(0xffffffffu, take);
      if (lane >= o && ((takes >> (lane - o)) & 1u)) row = -1;  // absorbed by the head at lane - o
    }
  }
  if (row >= 0) atomicAdd(acc_hi + row, acc);
}

template <typename T, bool WEIGHTED>
__global__ void __launch_bounds__(kHotThreads, 1)
k_spmv_blocked(hot_unit_t const* __restrict__ units, int n_units, int* __restrict__ unit_counter,
               int32_t const* __restrict__ slot_row, uint16_t const* __restrict__ idx16,
               int32_t const* __restrict__ idx32, long long cold_slot0, T const* __restrict__ w,
               T const* __restrict__ x, double* __restrict__ acc_hi, int W, int B, pr_state_t const* __restrict__ st)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* sx = reinterpret_cast<T*>(smem_raw);
  __shared__ uint64_t bar;
  __shared__ int s_next;
  if (st->done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    s_next = atomicAdd(unit_counter, 1);
  }
  if (threadIdx.x < kHotZeroPad) sx[W + threadIdx.x] = (T)0;  // the padding column(s) of every slice
  unsigned phase = 0;
  int cur_block  = -1;
  while (true) {
    __syncthreads();  // s_next is published; everyone is done with the previous unit's slice
    const int u = s_next;
    __syncthreads();
    if (u >= n_units) break;
    if (threadIdx.x == 0) s_next = atomicAdd(unit_counter, 1);  // fetch the next unit while working
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
    // two warp steps (64 slots) per iteration, software pipelined: the index vectors and rows of the next
    // iteration are loaded before the current one is gathered and reduced
    const int stride = kHotWarps * 64;
    int s0           = un.slot_begin + warp * 64;
    slot_ids_t ia, ib;
    int ra = -1, rb = -1;
    {
      const int sa = s0 + lane, sb = s0 + 32 + lane;
      const bool va = sa < un.slot_end, vb = sb < un.slot_end;
      if (hot) { ia = hot_slot_load<true>(sa, va, idx16, idx32, cold_slot0); ib = hot_slot_load<true>(sb, vb, idx16, idx32, cold_slot0); }
      else { ia = hot_slot_load<false>(sa, va, idx16, idx32, cold_slot0); ib = hot_slot_load<false>(sb, vb, idx16, idx32, cold_slot0); }
      ra = va ? slot_row[sa] : -1;
      rb = vb ? slot_row[sb] : -1;
    }
    for (; s0 < un.slot_end; s0 += stride) {
      const int sa = s0 + lane, sb = s0 + 32 + lane;
      const bool va = sa < un.slot_end, vb = sb < un.slot_end;
      // prefetch the next iteration
      const int na = sa + stride, nb = sb + stride;
      const bool nva = na < un.slot_end, nvb = nb < un.slot_end;
      slot_ids_t ja, jb;
      if (hot) { ja = hot_slot_load<true>(na, nva, idx16, idx32, cold_slot0); jb = hot_slot_load<true>(nb, nvb, idx16, idx32, cold_slot0); }
      else { ja = hot_slot_load<false>(na, nva, idx16, idx32, cold_slot0); jb = hot_slot_load<false>(nb, nvb, idx16, idx32, cold_slot0); }
      const int nra = nva ? slot_row[na] : -1;
      const int nrb = nvb ? slot_row[nb] : -1;
      double aa, ab;
      if (hot) {
        aa = hot_slot_sum<T, WEIGHTED, true>(ia, sa, va, w, x, sx);
        ab = hot_slot_sum<T, WEIGHTED, true>(ib, sb, vb, w, x, sx);
      } else {
        aa = hot_slot_sum<T, WEIGHTED, false>(ia, sa, va, w, x, sx);
        ab = hot_slot_sum<T, WEIGHTED, false>(ib, sb, vb, w, x, sx);
      }
      if (un.combine) {
        hot_emit<true>(aa, ra, acc_hi, lane);
        if (s0 + 32 < un.slot_end) hot_emit<true>(ab, rb, acc_hi, lane);
      } else {
        hot_emit<false>(aa, ra, acc_hi, lane);
        if (s0 + 32 < un.slot_end) hot_emit<false>(ab, rb, acc_hi, lane);
      }
      ia = ja; ib = jb; ra = nra; rb = nrb;
    }
  }
}

// y[row] = acc * alpha + init for every degree>=32 row; clears the accumulators and the unit counter
template <typename T>
__global__ void k_spmv_blocked_finish(double* __restrict__ acc_hi, int n_hi, T* __restrict__ y,
                                      int32_t const* __restrict__ row_vertex, double alpha, int* unit_counter,
                                      pr_state_t const* __restrict__ st)
{
  if (st->done) return;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r == 0) *unit_counter = 0;
  if (r >= n_hi) return;
  y[row_vertex ? row_vertex[r] : r] = (T)(acc_hi[r] * alpha + st->init);
  acc_hi[r]                          = 0.0;
}

template <typename O, typename T>
void launch_low_rows(handle_impl const& h, csx_t const& c, T const* x, T* y, double alpha, pr_state_t const* st)
{
  low_bins_t bins = make_low_bins(c);
  int lblocks     = bins.block_begin[kNumSeg - 1];
  if (lblocks <= 0) return;
  if (c.weights.data())
    B200_LAUNCH(h, (k_spmv_low<O, T, true>), lblocks, 256, 0, c.offsets.as<O>(), c.indices.as<int32_t>(),
                c.weights.as<T>(), x, y, c.row_vertex.as<int32_t>(), bins, alpha, st);
  else
    B200_LAUNCH(h, (k_spmv_low<O, T, false>), lblocks, 256, 0, c.offsets.as<O>(), c.indices.as<int32_t>(),
                c.weights.as<T>(), x, y, c.row_vertex.as<int32_t>(), bins, alpha, st);
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
  const int grid = std::min(L.n_cta, L.n_units);
  if (L.slot_w.data())
    B200_LAUNCH(h, (k_spmv_blocked<T, true>), grid, kHotThreads, kHotDynSmem, L.units.as<hot_unit_t>(), L.n_units,
                L.unit_counter.as<int>(), L.slot_row.as<int32_t>(), L.slot_idx16.as<uint16_t>(), L.slot_idx32.as<int32_t>(),
                (long long)L.n_hot_slots, L.slot_w.as<T>(), x, acc_hi, L.W, L.B, st);
  else
    B200_LAUNCH(h, (k_spmv_blocked<T, false>), grid, kHotThreads, kHotDynSmem, L.units.as<hot_unit_t>(), L.n_units,
                L.unit_counter.as<int>(), L.slot_row.as<int32_t>(), L.slot_idx16.as<uint16_t>(), L.slot_idx32.as<int32_t>(),
                (long long)L.n_hot_slots, L.slot_w.as<T>(), x, acc_hi, L.W, L.B, st);
  B200_LAUNCH(h, (k_spmv_blocked_finish<T>), (L.n_hi + 255) / 256, 256, 0, acc_hi, L.n_hi, y, c.row_vertex.as<int32_t>(),
              alpha, L.unit_counter.as<int>(), st);
  launch_low_rows<O, T>(h, c, x, y, alpha, st);
}

// dispatch: blocked layout when it exists for this graph, else the plain edge-balanced sweep
template <typename O, typename T>
void launch_pull_sweep_auto(handle_impl const& h, csx_t const& c, int32_t n_vertices, T const* x, T* y, double* acc_hi,
                            double alpha, pr_state_t const* st)
{
  hot_layout_t const* L = hot_layout(h, c, n_vertices, sizeof(T));
  if (L) launch_pull_sweep_blocked<O, T>(h, c, *L, x, y, acc_hi, alpha, st);
  else launch_pull_sweep<O, T>(h, c, x, y, acc_hi, alpha, st);
}

// elements an x buffer needs: whole slices are TMA-copied and x[n_vertices] must be a readable zero.
// The buffer must be zero-filled once at allocation; only [0, n_vertices) is ever written afterwards.
inline size_t padded_x_elems(int32_t n_vertices, size_t elem_size)
{
  const size_t slice = kHotSliceBytes / elem_size;
  const size_t W     = slice - kHotZeroPad;
  return ((size_t)n_vertices / W + 2) * slice;
}

}  // namespace b200
