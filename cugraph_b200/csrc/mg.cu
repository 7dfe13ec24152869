// Multi-GPU support (placeholder until the 2D-partitioned path lands in this file).
#include "graph.cuh"

namespace b200 {

void attach_comm(handle_impl*, void*)
{
  throw capi_exception(CUGRAPH_NOT_IMPLEMENTED, "multi-GPU handles are not available in this build");
}

void free_mg_graph(graph_impl*) {}

void mg_pagerank(handle_impl const&, graph_impl&, mg_pr_args const&, centrality_result_impl&)
{
  throw capi_exception(CUGRAPH_NOT_IMPLEMENTED, "multi-GPU PageRank is not available in this build");
}

}  // namespace b200

using namespace b200;
extern "C" {
cugraph_error_code_t cugraph_b200_get_nccl_unique_id(byte_t*, cugraph_error_t** error)
{ return guarded(error, [&] { throw capi_exception(CUGRAPH_NOT_IMPLEMENTED, "nccl"); }); }
cugraph_error_code_t cugraph_b200_comm_create(const byte_t*, int, int, cugraph_b200_comm_t**, cugraph_error_t** error)
{ return guarded(error, [&] { throw capi_exception(CUGRAPH_NOT_IMPLEMENTED, "nccl"); }); }
void cugraph_b200_comm_free(cugraph_b200_comm_t*) {}
cugraph_error_code_t cugraph_graph_create_mg(cugraph_resource_handle_t const*, cugraph_graph_properties_t const*,
  cugraph_type_erased_device_array_view_t const* const*, cugraph_type_erased_device_array_view_t const* const*,
  cugraph_type_erased_device_array_view_t const* const*, cugraph_type_erased_device_array_view_t const* const*,
  cugraph_type_erased_device_array_view_t const* const*, cugraph_type_erased_device_array_view_t const* const*,
  bool_t, size_t, bool_t, bool_t, bool_t, bool_t, cugraph_graph_t**, cugraph_error_t** error)
{ return guarded(error, [&] { throw capi_exception(CUGRAPH_NOT_IMPLEMENTED, "mg"); }); }
cugraph_error_code_t cugraph_graph_create_with_times_mg(cugraph_resource_handle_t const*, cugraph_graph_properties_t const*,
  cugraph_type_erased_device_array_view_t const* const*, cugraph_type_erased_device_array_view_t const* const*,
  cugraph_type_erased_device_array_view_t const* const*, cugraph_type_erased_device_array_view_t const* const*,
  cugraph_type_erased_device_array_view_t const* const*, cugraph_type_erased_device_array_view_t const* const*,
  cugraph_type_erased_device_array_view_t const* const*, cugraph_type_erased_device_array_view_t const* const*,
  bool_t, size_t, bool_t, bool_t, bool_t, bool_t, cugraph_graph_t**, cugraph_error_t** error)
{ return guarded(error, [&] { throw capi_exception(CUGRAPH_NOT_IMPLEMENTED, "mg"); }); }
}
