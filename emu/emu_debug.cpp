// Accessors for the CPU emulation build (libcugraph_c_emu.so): test infrastructure only, see emu/cuda_runtime.h.
#include "graph.cuh"
#include "spmv_hot_x.cuh"

#include <algorithm>
#include <vector>

namespace b200 {
alignas(128) unsigned char smem_raw[256 * 1024];  // the dynamic shared memory of the one CTA that runs at a time
}  // namespace b200

using namespace b200;

#define EMU_EXPORT extern "C" __attribute__((visibility("default")))

// primary orientation: pointers into the ("device" = host) arrays
EMU_EXPORT int emu_graph_primary(cugraph_graph_t* graph, int64_t* ints /*[8]: n_rows,nnz,offs64,nnz_hi,n_vertices,weighted,wsize,0*/,
                                 int32_t* seg /*[8]*/, void** ptrs /*[5]: offsets,indices,weights,ext_of_int,row_vertex*/)
{
  auto* g        = reinterpret_cast<graph_impl*>(graph);
  csx_t const& c = *g->primary;
  ints[0] = c.n_rows; ints[1] = c.nnz; ints[2] = c.offs64; ints[3] = c.nnz_hi; ints[4] = g->n_vertices;
  ints[5] = g->weighted; ints[6] = (int64_t)dtype_size(g->weight_type); ints[7] = 0;
  for (int k = 0; k <= kNumSeg; ++k) seg[k] = c.seg[k];
  ptrs[0] = c.offsets.data(); ptrs[1] = c.indices.data(); ptrs[2] = c.weights.data(); ptrs[3] = g->ext_of_int.data();
  ptrs[4] = c.row_vertex.data();
  return 0;
}

// column-blocked piece layout of the primary orientation (built on first use; honours the CUGRAPH_B200_HOT_* switches)
EMU_EXPORT int emu_hot_layout(const cugraph_resource_handle_t* handle, cugraph_graph_t* graph, int64_t* ints /*[12]*/,
                              void** ptrs /*[10]*/)
{
  auto const& h  = H(handle);
  auto* g        = reinterpret_cast<graph_impl*>(graph);
  csx_t const& c = *g->primary;
  const size_t es = g->weighted ? dtype_size(g->weight_type) : 4;
  hot_layout_t const* L = nullptr;
  try {
    L = hot_layout(h, c, g->n_vertices, es);
  } catch (std::exception const& e) {
    std::fprintf(stderr, "emu_hot_layout: %s\n", e.what());
    return 2;
  }
  if (!L) return 1;
  ints[0] = L->W; ints[1] = L->B; ints[2] = L->n_hi; ints[3] = L->nnz_hi; ints[4] = L->n_hot_slots; ints[5] = L->n_slots;
  ints[6] = L->n_subs; ints[7] = L->n_units; ints[8] = L->n_cta; ints[9] = L->narrow; ints[10] = (int64_t)es; ints[11] = 0;
  ptrs[0] = L->slot_idx16.data(); ptrs[1] = L->slot_idx32.data(); ptrs[2] = L->slot_w.data(); ptrs[3] = L->seg_row.data();
  ptrs[4] = L->subs.data(); ptrs[5] = L->units.data(); ptrs[6] = L->cta_range.data(); ptrs[7] = L->slot_idx_h.data();
  ptrs[8] = L->slot_idx_q.data(); ptrs[9] = L->slot_idx_s.data();
  return 0;
}

// y[low rows] = alpha * sum x[src] * w + init through the exact-degree ELL copy and k_spmv_low_ell (float graphs)
EMU_EXPORT int emu_low_ell_sweep(const cugraph_resource_handle_t* handle, cugraph_graph_t* graph, const float* x, float* y,
                                 double alpha, double init)
{
  auto const& h  = H(handle);
  auto* g        = reinterpret_cast<graph_impl*>(graph);
  csx_t const& c = *g->primary;
  if (g->weighted && g->weight_type != FLOAT32) return 3;
  low_ell_t const* E = nullptr;
  try {
    E = low_ell_layout(h, c, 4);
  } catch (std::exception const& e) {
    std::fprintf(stderr, "emu_low_ell_sweep: %s\n", e.what());
    return 2;
  }
  if (!E) return 1;
  pr_state_t st{};
  st.init = init;
  launch_low_rows_ell<float>(h, c, *E, x, y, alpha, &st);
  return 0;
}

// Functional model of k_spmv_blocked / k_spmv_blocked_x on the piece layout: the kernels' own per-lane device functions
// (hot_run_groups, hot_run_groups_c1, hot_run_groups_narrow, hot_slot_sum, hot_emit) called lane by lane, units in order,
// groups dealt to the 32 warps exactly as in the kernels; the shared-memory slice is a host copy of x[b*W, b*W+W) + zeros.
// acc[row] receives the fp64 sums of the degree >= 32 rows.  mode bit 0: use the four-groups-in-flight path for the
// one-slot class (k_spmv_blocked_x); narrow classes are always routed as in k_spmv_blocked_x.
template <typename T, bool WEIGHTED>
static void model_blocked(hot_layout_t const& L, T const* x, double* acc, int mode)
{
  auto const* units = L.units.as<hot_unit_t>();
  auto const* subs  = L.subs.as<hot_sub_t>();
  auto const* seg_row = L.seg_row.as<int32_t>();
  auto const* idx16 = L.slot_idx16.as<uint16_t>();
  auto const* idx32 = L.slot_idx32.as<int32_t>();
  auto const* idx_h = L.slot_idx_h.as<uint2>();
  auto const* idx_q = L.slot_idx_q.as<uint32_t>();
  auto const* idx_s = L.slot_idx_s.as<uint16_t>();
  T const* w        = L.slot_w.as<T>();
  const int cold0   = (int)L.n_hot_slots;
  std::vector<T> sx((size_t)L.W + kHotZeroPad);
  int cur_block = -1;
  for (int u = 0; u < L.n_units; ++u) {
    const hot_unit_t un = units[u];
    const bool hot      = un.block < L.B;
    if (hot && un.block != cur_block) {
      std::copy(x + (size_t)un.block * L.W, x + (size_t)un.block * L.W + L.W, sx.begin());
      std::fill(sx.begin() + L.W, sx.end(), (T)0);
      cur_block = un.block;
    }
    int dealt = 0;
    for (int si = un.sub_begin; si < un.sub_end; ++si) {
      const hot_sub_t sb = subs[si];
      for (int warp = 0; warp < kHotWarps; ++warp) {
        const int q0 = (warp - dealt) & (kHotWarps - 1);
        for (int lane = 0; lane < 32; ++lane) {
          if (sb.cls > 8) {
            if (sb.cls == 16) hot_run_groups_narrow<T, 4>(sb, q0, lane, seg_row, idx_h, idx_q, idx_s, sx.data(), acc);
            else if (sb.cls == 32) hot_run_groups_narrow<T, 2>(sb, q0, lane, seg_row, idx_h, idx_q, idx_s, sx.data(), acc);
            else hot_run_groups_narrow<T, 1>(sb, q0, lane, seg_row, idx_h, idx_q, idx_s, sx.data(), acc);
          } else if ((mode & 1) && sb.cls == 1 && hot) {
            hot_run_groups_c1<T, WEIGHTED, true>(sb, q0, lane, seg_row, idx16, idx32, cold0, w, x, sx.data(), acc);
          } else if (hot) {
            hot_run_groups<T, WEIGHTED, true>(sb, q0, lane, seg_row, idx16, idx32, cold0, w, x, sx.data(), acc);
          } else {
            hot_run_groups<T, WEIGHTED, false>(sb, q0, lane, seg_row, idx16, idx32, cold0, w, x, sx.data(), acc);
          }
        }
      }
      dealt += sb.n_groups;
    }
  }
}

EMU_EXPORT int emu_blocked_sweep(const cugraph_resource_handle_t* handle, cugraph_graph_t* graph, const float* x /* padded */,
                                 double* acc /* n_hi */, int mode)
{
  auto const& h  = H(handle);
  auto* g        = reinterpret_cast<graph_impl*>(graph);
  csx_t const& c = *g->primary;
  if (g->weighted && g->weight_type != FLOAT32) return 3;
  hot_layout_t const* L = nullptr;
  try {
    L = hot_layout(h, c, g->n_vertices, 4);
  } catch (std::exception const& e) {
    std::fprintf(stderr, "emu_blocked_sweep: %s\n", e.what());
    return 2;
  }
  if (!L) return 1;
  if (L->slot_w.data()) model_blocked<float, true>(*L, x, acc, mode);
  else model_blocked<float, false>(*L, x, acc, mode);
  return 0;
}

EMU_EXPORT size_t emu_padded_x_elems(int32_t nv, size_t es) { return padded_x_elems(nv, es); }

// model of k_spmv_low_ell_hot (spmv_hot.cuh): `grid` persistent CTAs of four virtual 256-thread blocks, gathers of the first
// W sources served from a copy of x[0, W) (the kernel's shared-memory slice), the rest from x
EMU_EXPORT int emu_low_ell_hot_sweep(const cugraph_resource_handle_t* handle, cugraph_graph_t* graph, const float* x /* padded */,
                                     float* y, double alpha, double init, int grid)
{
  auto const& h  = H(handle);
  auto* g        = reinterpret_cast<graph_impl*>(graph);
  csx_t const& c = *g->primary;
  if (g->weighted && g->weight_type != FLOAT32) return 3;
  low_ell_t const* E = nullptr;
  try {
    E = low_ell_layout(h, c, 4);
  } catch (std::exception const& e) {
    std::fprintf(stderr, "emu_low_ell_hot_sweep: %s\n", e.what());
    return 2;
  }
  if (!E) return 1;
  const int W = (int)(kHotSliceBytes / sizeof(float)) - kHotZeroPad;
  std::vector<float> sx(x, x + W);
  gather_hot_t<float> gh{x, sx.data(), W};
  low_ell_args_t a   = make_low_ell_args(*E);
  const int n_vblock = a.block_begin[32];
  for (int cta = 0; cta < grid; ++cta)
    for (int sub = 0; sub < 4; ++sub)
      for (int vb = cta * 4 + sub; vb < n_vblock; vb += grid * 4)
        for (int vtid = 0; vtid < 256; ++vtid) {
          if (E->w.data())
            low_ell_block<float, true>(vb, vtid, E->idx.as<int32_t>(), E->w.as<float>(), gh, y, c.row_vertex.as<int32_t>(), a, alpha, init);
          else
            low_ell_block<float, false>(vb, vtid, E->idx.as<int32_t>(), E->w.as<float>(), gh, y, c.row_vertex.as<int32_t>(), a, alpha, init);
        }
  return 0;
}

// forget the cached layouts of the primary orientation (so that another set of CUGRAPH_B200_* switches can be staged)
EMU_EXPORT void emu_reset_layouts(cugraph_graph_t* graph)
{
  auto* g        = reinterpret_cast<graph_impl*>(graph);
  csx_t const& c = *g->primary;
  c.hot4.reset();
  c.hot8.reset();
  c.hot4_tried = c.hot8_tried = false;
  c.low_ell.reset();
  c.low_ell_tried = false;
}
