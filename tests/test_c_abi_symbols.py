"""CPU-only: the C-ABI library builds/loads without a GPU and exports every symbol declared in
include/cugraph_c/*.h (no compute call is made here)."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for path in glob.glob(os.path.join(ROOT, "include", "cugraph_c", "*.h")):
        text = open(path).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"CUGRAPH_EXPORT\s+[^;(]*?\b(cugraph_\w+)\s*\(", text, flags=re.S):
            names.add(m.group(1))
    return names


def _lib():
    from cugraph_b200 import build
    path = build.build()
    return ctypes.CDLL(path)


def test_every_declared_symbol_is_exported():
    lib = _lib()
    names = _declared()
    assert len(names) >= 50
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, f"declared but not exported: {missing}"


def test_hot_path_symbols_present():
    names = _declared()
    for n in ["cugraph_pagerank", "cugraph_pagerank_allow_nonconvergence", "cugraph_personalized_pagerank",
              "cugraph_bfs", "cugraph_sssp", "cugraph_graph_create_sg", "cugraph_graph_create_with_times_sg",
              "cugraph_graph_create_sg_from_csr", "cugraph_create_resource_handle",
              "cugraph_type_erased_device_array_view_create", "cugraph_centrality_result_get_values",
              "cugraph_paths_result_get_distances"]:
        assert n in names


def test_python_binding_binds_all(monkeypatch):
    from cugraph_b200 import _capi
    L = _capi.lib()
    assert L.cugraph_b200_version().startswith(b"cugraph_b200")


def test_error_object_roundtrip_without_gpu():
    """Creating a view and reading it back needs no device."""
    from cugraph_b200 import _capi
    L = _capi.lib()
    v = L.cugraph_type_erased_device_array_view_create(ctypes.c_void_p(0x1000), 7, _capi.FLOAT32)
    assert L.cugraph_type_erased_device_array_view_size(v) == 7
    assert L.cugraph_type_erased_device_array_view_type(v) == _capi.FLOAT32
    L.cugraph_type_erased_device_array_view_free(v)
    assert L.cugraph_error_message(None) is None
