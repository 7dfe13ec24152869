// Column-blocked pull sweep for the degree>=32 rows: gathers served from SHARED MEMORY.
//
// Why: on RMAT-24 the plain sweep (spmv.cuh) is bound by the L2 -> SM path, not by HBM: every gather
// of x[src] costs a 32-byte L2 sector for 4 useful bytes (ncu: 0.86 L2 sectors per edge, lts 67 %,
// l1tex 73 %, DRAM 16 % — profiles/r01_ncu_k_spmv_hi_v1.csv).  The source space is therefore cut into
// B hot blocks of W vertices whose x-slice (192 KiB) a persistent CTA keeps in shared memory, filled by
// TMA bulk copies (cp.async.bulk + mbarrier).  Rows keep their neighbours sorted by source id, so a
// row's adjacency is already partitioned by block; staging stores the (row, block) segments block-major
// with 16-bit local column ids plus, for every non-empty segment, its start and its row
// (hot_layout_t, graph.cuh).
//
// Execution: work units (128 chunks of 1024 edges, all of one block) are handed out dynamically
// through one atomic counter (the next unit is fetched while the current one is processed); a CTA
// refills its shared memory only when its next unit belongs to another block.  A warp owns a chunk;
// g = 1..32 lanes cooperate on one segment (g from the chunk's average segment length:
// vertex-group-per-warp), each lane keeps four gathers in flight, partial sums are folded with log2(g)
// shuffles and ONE fp64 atomic per segment piece goes to acc_hi[row].  The cold block
// (sources >= B*W) runs through the same code with global gathers.
#pragma once
#include "spmv.cuh"

namespace b200 {

constexpr bool kHotStageTiles = false;  // TMA-staged index tiles measured slower (16 warps): r01 notes
constexpr int kHotThreads   = kHotStageTiles ? 512 : 1024;
constexpr int kHotWarps     = kHotThreads / 32;
constexpr int kHotChunk     = 1024;
constexpr int kHotSmemBytes = kHotSliceBytes;  // x slice (graph.cuh)
constexpr int kHotTileBytes = kHotChunk * 2;   // one chunk of 16-bit column ids
constexpr int kHotDynSmem   = kHotSmemBytes + (kHotStageTiles ? kHotWarps * 2 * kHotTileBytes : 0);  // slice (+ index tiles)
constexpr int kHotTmaPiece  = 16 * 1024;       // bytes per bulk copy of the slice

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

// One chunk = up to 1024 consecutive edges of one block = n_seg (row, block) segments.
// HOT: gather from the shared-memory slice; else (cold block) from global x.
template <typename T, bool WEIGHTED, bool HOT>
__device__ __forceinline__ void hot_process_chunk(int cb, int ce, int seg0, int n_seg, int lg,
                                                  int32_t const* __restrict__ seg_start,
                                                  int32_t const* __restrict__ seg_row,
                                                  uint16_t const* __restrict__ tile /* smem, index 0 = edge cb */,
                                                  int32_t const* __restrict__ idx32,
                                                  int cold_base, T const* __restrict__ w, T const* __restrict__ x,
                                                  T const* __restrict__ sx, double* __restrict__ acc_hi, int lane)
{
  const uint16_t* idx16 = tile;  // idx16[i] is the column of permuted position i
  const int g      = 1 << lg;
  const int sub    = lane & (g - 1);
  const int groups = 32 >> lg;
  int j            = lane >> lg;
  // bounds of this group's first segment (afterwards prefetched one pass ahead)
  int lo = ce, hi = ce, row = 0;
  if (j < n_seg) {
    lo  = seg_start[seg0 + j];
    hi  = seg_start[seg0 + j + 1];
    row = seg_row[seg0 + j];
  }
  while (__any_sync(0xffffffffu, j < n_seg)) {
    const int jn = j + groups;
    int lo_n = ce, hi_n = ce, row_n = 0;
    if (jn < n_seg) {
      lo_n  = seg_start[seg0 + jn];
      hi_n  = seg_start[seg0 + jn + 1];
      row_n = seg_row[seg0 + jn];
    }
    lo = lo < cb ? cb : lo;
    hi = hi > ce ? ce : hi;
    double acc = 0.0;
    // rounds of kR predicated edges per lane: all index loads of a round are issued back to back
    // (no serial remainder loop: every load of the round is in flight together)
    constexpr int kR = 8;
    for (int i = lo + sub; i < hi; i += kR * g) {
      unsigned c[kR];
      T wv[kR];
#pragma unroll
      for (int k = 0; k < kR; ++k) {
        const int e = i + k * g;
        c[k]        = 0;
        wv[k]       = (T)0;
        if (e < hi) {
          c[k]  = HOT ? (unsigned)idx16[e] : (unsigned)idx32[e - cold_base];
          wv[k] = WEIGHTED ? w[e] : (T)1;
        }
      }
      T v[kR];
#pragma unroll
      for (int k = 0; k < kR; ++k) v[k] = (HOT ? sx[c[k]] : x[c[k]]) * wv[k];
      double part = 0.0;
#pragma unroll
      for (int k = 0; k < kR; k += 4) part += ((double)v[k] + (double)v[k + 1]) + ((double)v[k + 2] + (double)v[k + 3]);
      acc += part;
    }
    for (int o = g >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (sub == 0 && hi > lo) atomicAdd(acc_hi + row, acc);
    j   = jn;
    lo  = lo_n;
    hi  = hi_n;
    row = row_n;
  }
}

// a work unit: up to 128 consecutive chunks of one block (mirrored by hot_unit_host_t, graph_build.cu)
struct hot_unit_t {
  int32_t chunk_begin;
  int32_t chunk_end;
  int32_t block;
  int32_t pos_begin;   // permuted position of the first chunk's first edge
  int32_t block_end;   // permuted position one past the block's last edge
  int32_t head_begin;  // (unused by the kernel) head word of the first chunk
  int32_t pad0, pad1;
};

template <typename T, bool WEIGHTED>
__global__ void __launch_bounds__(kHotThreads, 1)
k_spmv_blocked(hot_unit_t const* __restrict__ units, int n_units, int* __restrict__ unit_counter,
               int32_t const* __restrict__ chunk_seg0, int cold_base, int32_t const* __restrict__ seg_start,
               int32_t const* __restrict__ seg_row, uint16_t const* __restrict__ idx16,
               int32_t const* __restrict__ idx32, T const* __restrict__ w, T const* __restrict__ x,
               double* __restrict__ acc_hi, int W, int B, pr_state_t const* __restrict__ st)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* sx = reinterpret_cast<T*>(smem_raw);
  __shared__ uint64_t bar;                      // slice fill
  __shared__ uint64_t tile_bar[kHotWarps][2];   // per-warp index tiles, double buffered
  __shared__ int s_next;
  if (st->done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint16_t* my_tiles = reinterpret_cast<uint16_t*>(smem_raw + kHotSmemBytes + warp * 2 * kHotTileBytes);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < kHotWarps; ++i) {
      mbar_init(&tile_bar[i][0], 1);
      mbar_init(&tile_bar[i][1], 1);
    }
    s_next = atomicAdd(unit_counter, 1);
  }
  unsigned phase = 0, tphase0 = 0, tphase1 = 0;
  int cur_block = -1;
  while (true) {
    __syncthreads();  // s_next is published; everyone is done with the previous unit's slice
    const int u = s_next;
    __syncthreads();
    if (u >= n_units) break;
    if (threadIdx.x == 0) s_next = atomicAdd(unit_counter, 1);  // fetch the next unit while working
    const hot_unit_t un = units[u];
    const int b         = un.block;
    const bool hot      = b < B;
    int k               = un.chunk_begin + warp;
    // first index tile of this warp: in flight while the slice is (re)filled
    int buf = 0;
    if (kHotStageTiles && hot && k < un.chunk_end && lane == 0) {
      const int cb = un.pos_begin + (k - un.chunk_begin) * kHotChunk;
      const int ce = (cb + kHotChunk < un.block_end) ? cb + kHotChunk : un.block_end;
      const unsigned bytes = (unsigned)(((ce - cb) * 2 + 15) & ~15);
      mbar_expect_tx(&tile_bar[warp][0], bytes);
      tma_bulk_g2s(my_tiles, idx16 + cb, bytes, &tile_bar[warp][0]);
    }
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
    int seg0 = 0, seg1 = 0;
    if (k < un.chunk_end) {
      seg0 = chunk_seg0[k];
      seg1 = chunk_seg0[k + 1];
    }
    while (k < un.chunk_end) {
      const int cb = un.pos_begin + (k - un.chunk_begin) * kHotChunk;
      const int ce = (cb + kHotChunk < un.block_end) ? cb + kHotChunk : un.block_end;
      const int kn  = k + kHotWarps;  // next chunk of this warp: prefetch its metadata and its index tile
      int seg0_next = 0, seg1_next = 0;
      if (kn < un.chunk_end) {
        seg0_next = chunk_seg0[kn];
        seg1_next = chunk_seg0[kn + 1];
        if (kHotStageTiles && hot) {
          __syncwarp();  // every lane is done reading the tile that is about to be overwritten
          if (lane == 0) {
            const int cbn = un.pos_begin + (kn - un.chunk_begin) * kHotChunk;
            const int cen = (cbn + kHotChunk < un.block_end) ? cbn + kHotChunk : un.block_end;
            const unsigned bytes = (unsigned)(((cen - cbn) * 2 + 15) & ~15);
            mbar_expect_tx(&tile_bar[warp][buf ^ 1], bytes);
            tma_bulk_g2s(my_tiles + (buf ^ 1) * kHotChunk, idx16 + cbn, bytes, &tile_bar[warp][buf ^ 1]);
          }
        }
      }
      const int n_seg = seg1 - seg0 + 1;  // segments overlapping the chunk (the last may be empty here)
      const int avg   = (ce - cb) / n_seg;
      int lg          = 0;  // lanes per segment: 4-8 edges per lane (8-16 measured slower)
      while (lg < 5 && (8 << lg) <= avg) ++lg;
      if (hot) {
        const uint16_t* tile = idx16;
        if (kHotStageTiles) {
          if (buf == 0) {
            mbar_wait(&tile_bar[warp][0], tphase0);
            tphase0 ^= 1;
          } else {
            mbar_wait(&tile_bar[warp][1], tphase1);
            tphase1 ^= 1;
          }
          tile = my_tiles + buf * kHotChunk - cb;
          buf ^= 1;
        }
        hot_process_chunk<T, WEIGHTED, true>(cb, ce, seg0, n_seg, lg, seg_start, seg_row, tile, idx32, cold_base, w, x, sx,
                                             acc_hi, lane);
      } else {
        hot_process_chunk<T, WEIGHTED, false>(cb, ce, seg0, n_seg, lg, seg_start, seg_row, nullptr, idx32, cold_base, w, x,
                                              sx, acc_hi, lane);
      }
      k    = kn;
      seg0 = seg0_next;
      seg1 = seg1_next;
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

// x must be readable up to roundup(n_vertices, W) elements (the TMA fill copies whole slices)
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
  if (L.w.data())
    B200_LAUNCH(h, (k_spmv_blocked<T, true>), grid, kHotThreads, kHotDynSmem, L.units.as<hot_unit_t>(), L.n_units,
                L.unit_counter.as<int>(), L.chunks.as<int32_t>(), (int)L.nnz_hot, L.seg_start.as<int32_t>(),
                L.seg_row.as<int32_t>(), L.idx16.as<uint16_t>(), L.idx32.as<int32_t>(), L.w.as<T>(), x, acc_hi, L.W, L.B, st);
  else
    B200_LAUNCH(h, (k_spmv_blocked<T, false>), grid, kHotThreads, kHotDynSmem, L.units.as<hot_unit_t>(), L.n_units,
                L.unit_counter.as<int>(), L.chunks.as<int32_t>(), (int)L.nnz_hot, L.seg_start.as<int32_t>(),
                L.seg_row.as<int32_t>(), L.idx16.as<uint16_t>(), L.idx32.as<int32_t>(), L.w.as<T>(), x, acc_hi, L.W, L.B, st);
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

// number of elements an x buffer needs so that whole shared-memory slices can be copied
inline size_t padded_x_elems(int32_t n_vertices, size_t elem_size)
{
  size_t W = kHotSmemBytes / elem_size;
  return (((size_t)n_vertices + W - 1) / W) * W;
}

}  // namespace b200
