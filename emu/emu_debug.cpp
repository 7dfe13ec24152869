// Accessors for the CPU emulation build (libcugraph_c_emu.so): test infrastructure only, see emu/cuda_runtime.h.
#include "graph.cuh"
#include "spmv.cuh"

namespace b200 {
void free_mg_graph(graph_impl*) {}
void attach_comm(handle_impl*, void*) {}
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
  ptrs[8] = L->slot_idx_q.data(); ptrs[9] = nullptr;
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
