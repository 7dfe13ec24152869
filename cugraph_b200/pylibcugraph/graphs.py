"""SGGraph / MGGraph (reference python/pylibcugraph/pylibcugraph/graphs.pyx)."""
import ctypes as C

from cugraph_b200 import _capi
from cugraph_b200.pylibcugraph.graph_properties import GraphProperties
from cugraph_b200.pylibcugraph.resource_handle import ResourceHandle
from cugraph_b200.pylibcugraph.utils import View, assert_CAI_type


class _GPUGraph:
    def __init__(self):
        self._ptr = None
        self._lib = _capi.lib()

    @property
    def ptr(self):
        return self._ptr

    def __del__(self):
        try:
            if getattr(self, "_ptr", None):
                self._lib.cugraph_graph_free(self._ptr)
                self._ptr = None
        except Exception:
            pass


def _check_flag(v, name):
    if not isinstance(v, (int, bool)):
        raise TypeError(f"expected int or bool for {name}, got {type(v)}")


def _cai_dtype(a):
    import numpy as np
    return np.dtype(a.__cuda_array_interface__["typestr"])


def _widen(a):
    import torch
    if a is None:
        return None
    t = a if isinstance(a, torch.Tensor) else torch.as_tensor(a)
    return t.to(torch.int64).contiguous()


def _ensure_valid_dtypes(src, dst, vertices, edge_ids, t_start, t_end):
    import warnings
    vertex_args = [src, dst, vertices, edge_ids]
    if len({_cai_dtype(a) for a in vertex_args if a is not None}) > 1:
        warnings.warn("The graph requires 'src_or_offset_array', 'dst_or_index_array' "
                      "'vertices_array' and 'edge_id_array' to match. "
                      "Those will be widened to 64-bit.", UserWarning)
        src, dst, vertices, edge_ids = (_widen(a) for a in vertex_args)
    if len({_cai_dtype(a) for a in (t_start, t_end) if a is not None}) > 1:
        warnings.warn("The graph requires 'edge_start_time_array' and 'edge_end_time_array' "
                      "to match. Those will be widened to 64-bit.", UserWarning)
        t_start, t_end = _widen(t_start), _widen(t_end)
    return src, dst, vertices, edge_ids, t_start, t_end


class SGGraph(_GPUGraph):
    """Single-GPU graph; argument list of graphs.pyx:150-168 (COO via
    cugraph_graph_create_with_times_sg, CSR via cugraph_graph_create_sg_from_csr)."""

    def __init__(self, resource_handle, graph_properties, src_or_offset_array, dst_or_index_array,
                 weight_array=None, store_transposed=False, renumber=False, do_expensive_check=False,
                 edge_id_array=None, edge_type_array=None, edge_start_time_array=None,
                 edge_end_time_array=None, input_array_format="COO", vertices_array=None,
                 drop_self_loops=False, drop_multi_edges=False, symmetrize=False):
        super().__init__()
        if not isinstance(resource_handle, ResourceHandle):
            raise TypeError("resource_handle must be a ResourceHandle")
        if not isinstance(graph_properties, GraphProperties):
            raise TypeError("graph_properties must be a GraphProperties")
        _check_flag(store_transposed, "store_transposed")
        _check_flag(renumber, "renumber")
        _check_flag(do_expensive_check, "do_expensive_check")
        assert_CAI_type(src_or_offset_array, "src_or_offset_array")
        assert_CAI_type(dst_or_index_array, "dst_or_index_array")
        for a, nm in ((vertices_array, "vertices_array"), (weight_array, "weight_array"),
                      (edge_id_array, "edge_id_array"), (edge_type_array, "edge_type_array"),
                      (edge_start_time_array, "edge_start_time_array"), (edge_end_time_array, "edge_end_time_array")):
            assert_CAI_type(a, nm, True)
        # unequal id widths: warn and widen to 64 bits, as the reference does (utilities/api_tools.py:328-364 warns, the C layer
        # casts: c_api/graph_sg.cpp cast_vertex_t)
        (src_or_offset_array, dst_or_index_array, vertices_array, edge_id_array, edge_start_time_array,
         edge_end_time_array) = _ensure_valid_dtypes(src_or_offset_array, dst_or_index_array, vertices_array, edge_id_array,
                                                     edge_start_time_array, edge_end_time_array)
        views = [View(a) for a in (vertices_array, src_or_offset_array, dst_or_index_array, weight_array,
                                   edge_id_array, edge_type_array, edge_start_time_array, edge_end_time_array)]
        v, s, d, w, eid, ety, t0, t1 = [x.ptr for x in views]
        g = C.c_void_p()
        err = C.c_void_p()
        L = self._lib
        resource_handle.order_after_caller()
        if input_array_format == "COO":
            code = L.cugraph_graph_create_with_times_sg(
                resource_handle.ptr, C.byref(graph_properties.c), v, s, d, w, eid, ety, t0, t1,
                int(store_transposed), int(renumber), int(drop_self_loops), int(drop_multi_edges),
                int(symmetrize), int(do_expensive_check), C.byref(g), C.byref(err))
            where = "cugraph_graph_create_with_times_sg()"
        elif input_array_format == "CSR":
            code = L.cugraph_graph_create_sg_from_csr(
                resource_handle.ptr, C.byref(graph_properties.c), s, d, w, eid, ety,
                int(store_transposed), int(renumber), int(symmetrize), int(do_expensive_check),
                C.byref(g), C.byref(err))
            where = "cugraph_sg_graph_create_from_csr()"
        else:
            raise ValueError("invalid 'input_array_format'. Only 'COO' and 'CSR' format are supported.")
        for x in views:
            x.free()
        _capi.check(code, err, where)
        self._ptr = g.value
        self._handle = resource_handle
