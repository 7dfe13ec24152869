// BFS / SSSP (placeholder entry points; the frontier engine lands here).
#include "graph.cuh"
using namespace b200;
extern "C" {
cugraph_type_erased_device_array_view_t* cugraph_paths_result_get_vertices(cugraph_paths_result_t* r)
{ return reinterpret_cast<cugraph_type_erased_device_array_view_t*>(reinterpret_cast<paths_result_impl*>(r)->vertices->new_view()); }
cugraph_type_erased_device_array_view_t* cugraph_paths_result_get_distances(cugraph_paths_result_t* r)
{ return reinterpret_cast<cugraph_type_erased_device_array_view_t*>(reinterpret_cast<paths_result_impl*>(r)->distances->new_view()); }
cugraph_type_erased_device_array_view_t* cugraph_paths_result_get_predecessors(cugraph_paths_result_t* r)
{ return reinterpret_cast<cugraph_type_erased_device_array_view_t*>(reinterpret_cast<paths_result_impl*>(r)->predecessors->new_view()); }
void cugraph_paths_result_free(cugraph_paths_result_t* r)
{
  if (!r) return;
  auto* p = reinterpret_cast<paths_result_impl*>(r);
  delete p->vertices; delete p->distances; delete p->predecessors; delete p;
}
cugraph_error_code_t cugraph_bfs(const cugraph_resource_handle_t*, cugraph_graph_t*, cugraph_type_erased_device_array_view_t*,
                                 bool_t, size_t, bool_t, bool_t, cugraph_paths_result_t**, cugraph_error_t** error)
{ return guarded(error, [&] { throw capi_exception(CUGRAPH_NOT_IMPLEMENTED, "bfs"); }); }
cugraph_error_code_t cugraph_sssp(const cugraph_resource_handle_t*, cugraph_graph_t*, size_t, double, bool_t, bool_t,
                                  cugraph_paths_result_t**, cugraph_error_t** error)
{ return guarded(error, [&] { throw capi_exception(CUGRAPH_NOT_IMPLEMENTED, "sssp"); }); }
}
