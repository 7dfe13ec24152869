"""The reference's own golden vectors (cpp/tests/c_api/{pagerank,bfs,sssp}_test.c, pylibcugraph/tests/test_pagerank.py /
test_sssp.py — fixtures in tests/golden/reference_golden.json) replayed through the C ABI of the EMULATED library: the same
entry points, arguments and tolerances as the reference's tests, without a GPU.  (tests/test_pagerank_gpu.py and
tests/test_traversal_gpu.py replay them on the GPU.)"""
import ctypes as C

import numpy as np
import pytest

from tests.test_emu_algorithms_cpu import _paths, _view_to_np
from tests.test_emu_staging_cpu import FLOAT32, FLOAT64, INT32, Props, emu  # noqa: F401

FLT_MAX = float(np.finfo(np.float32).max)


def _view(L, a):
    t = {np.dtype(np.int32): INT32, np.dtype(np.float32): FLOAT32, np.dtype(np.float64): FLOAT64}[a.dtype]
    return C.c_void_p(L.cugraph_type_erased_device_array_view_create(a.ctypes.data, a.size, t))


def graph(L, src, dst, w, store_transposed=False, renumber=False, wdtype=np.float32, keep=None):
    s, d = np.asarray(src, np.int32), np.asarray(dst, np.int32)
    ww = None if w is None else np.asarray(w, wdtype)
    if keep is not None:
        keep.extend([s, d, ww])
    g, err = C.c_void_p(), C.c_void_p()
    code = L.cugraph_graph_create_with_times_sg(C.c_void_p(L.handle), C.byref(Props(0, 0)), None, _view(L, s), _view(L, d),
                                                None if ww is None else _view(L, ww), None, None, None, None,
                                                int(store_transposed), int(renumber), 0, 0, 0, 0, C.byref(g), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    return g


def by_vertex(verts, vals, n):
    out = np.zeros(n, dtype=vals.dtype)
    out[verts] = vals
    return out


def centrality(L, res):
    for f in ("cugraph_centrality_result_get_vertices", "cugraph_centrality_result_get_values"):
        getattr(L, f).restype = C.c_void_p
        getattr(L, f).argtypes = [C.c_void_p]
    L.cugraph_centrality_result_converged.restype = C.c_int
    L.cugraph_centrality_result_converged.argtypes = [C.c_void_p]
    L.cugraph_centrality_result_free.argtypes = [C.c_void_p]
    v = _view_to_np(L, L.cugraph_centrality_result_get_vertices(res))
    p = _view_to_np(L, L.cugraph_centrality_result_get_values(res))
    conv = bool(L.cugraph_centrality_result_converged(res))
    L.cugraph_centrality_result_free(res)
    return v, p, conv


@pytest.mark.parametrize("renumber", [False, True])
@pytest.mark.parametrize("store_transposed", [False, True])
@pytest.mark.parametrize("case", ["pagerank_6", "pagerank_6_nonconverged", "pagerank_4"])
def test_c_api_pagerank_goldens(emu, golden, case, store_transposed, renumber):  # noqa: F811
    c = golden["c_api"][case]
    g = graph(emu, c["src"], c["dst"], c["weights"], store_transposed, renumber)
    res, err = C.c_void_p(), C.c_void_p()
    code = emu.cugraph_pagerank_allow_nonconvergence(C.c_void_p(emu.handle), g, None, None, None, None, C.c_double(c["alpha"]),
                                                     C.c_double(c["epsilon"]), C.c_size_t(c["max_iterations"]), 0,
                                                     C.byref(res), C.byref(err))
    assert code == 0, emu.cugraph_error_message(err)
    verts, vals, conv = centrality(emu, res)
    assert vals.dtype == np.float32
    np.testing.assert_allclose(by_vertex(verts, vals, c["num_vertices"]), c["values"], rtol=c["rel_tol"])
    assert conv == ("nonconverged" not in case)
    emu.cugraph_graph_free(g)


def test_c_api_pagerank_reports_nonconvergence_as_error(emu, golden):  # noqa: F811
    """cugraph_pagerank (not the _allow_nonconvergence variant): error code + message AND the result object"""
    c = golden["c_api"]["pagerank_6_nonconverged"]
    g = graph(emu, c["src"], c["dst"], c["weights"], True, False)
    res, err = C.c_void_p(), C.c_void_p()
    code = emu.cugraph_pagerank(C.c_void_p(emu.handle), g, None, None, None, None, C.c_double(c["alpha"]), C.c_double(c["epsilon"]),
                                C.c_size_t(c["max_iterations"]), 0, C.byref(res), C.byref(err))
    assert code == 1 and b"failed to converge" in emu.cugraph_error_message(err)     # CUGRAPH_UNKNOWN_ERROR
    verts, vals, conv = centrality(emu, res)
    np.testing.assert_allclose(by_vertex(verts, vals, c["num_vertices"]), c["values"], rtol=c["rel_tol"])
    assert not conv
    emu.cugraph_graph_free(g)


@pytest.mark.parametrize("case", ["personalized_pagerank_4", "personalized_pagerank_4_nonconverged"])
def test_c_api_personalized_pagerank_goldens(emu, golden, case):  # noqa: F811
    c = golden["c_api"][case]
    g = graph(emu, c["src"], c["dst"], c["weights"], False, False)
    pv = np.asarray(c["personalization_vertices"], np.int32)
    pw = np.asarray(c["personalization_values"], np.float32)
    res, err = C.c_void_p(), C.c_void_p()
    code = emu.cugraph_personalized_pagerank_allow_nonconvergence(
        C.c_void_p(emu.handle), g, None, None, None, None, _view(emu, pv), _view(emu, pw), C.c_double(c["alpha"]),
        C.c_double(c["epsilon"]), C.c_size_t(c["max_iterations"]), 0, C.byref(res), C.byref(err))
    assert code == 0, emu.cugraph_error_message(err)
    verts, vals, conv = centrality(emu, res)
    np.testing.assert_allclose(by_vertex(verts, vals, 4), c["values"], rtol=c["rel_tol"])
    assert conv == ("nonconverged" not in case)
    emu.cugraph_graph_free(g)


@pytest.mark.parametrize("name", ["karate.csv", "dolphins.csv", "Simple_1", "Simple_2"])
def test_pylibcugraph_pagerank_goldens(emu, golden, name):  # noqa: F811
    d = golden["pylibcugraph"][name]
    p = d["pagerank"]
    g = graph(emu, d["src"], d["dst"], d["weights"], True, False)
    res, err = C.c_void_p(), C.c_void_p()
    code = emu.cugraph_pagerank_allow_nonconvergence(C.c_void_p(emu.handle), g, None, None, None, None, C.c_double(p["alpha"]),
                                                     C.c_double(p["epsilon"]), C.c_size_t(p["max_iterations"]), 0,
                                                     C.byref(res), C.byref(err))
    assert code == 0, emu.cugraph_error_message(err)
    verts, vals, conv = centrality(emu, res)
    assert conv and verts.dtype == np.int32 and vals.dtype == np.float32
    for v, x in zip(verts.tolist(), vals.tolist()):
        assert x == pytest.approx(p["values"][v], 1e-4)
    emu.cugraph_graph_free(g)


@pytest.mark.parametrize("renumber", [False, True])
@pytest.mark.parametrize("store_transposed", [False, True])
def test_c_api_bfs_golden(emu, golden, store_transposed, renumber):  # noqa: F811
    c = golden["c_api"]["bfs_6"]
    g = graph(emu, c["src"], c["dst"], c["weights"], store_transposed, renumber)
    srcs = np.asarray(c["sources"], np.int32)
    res, err = C.c_void_p(), C.c_void_p()
    code = emu.cugraph_bfs(C.c_void_p(emu.handle), g, _view(emu, srcs), 0, C.c_size_t(c["depth_limit"]), 1, 0, C.byref(res), C.byref(err))
    assert code == 0, emu.cugraph_error_message(err)
    verts, dist, pred = _paths(emu, res)
    assert by_vertex(verts, dist, 6).tolist() == c["distances"]
    assert by_vertex(verts, pred, 6).tolist() == c["predecessors"]
    emu.cugraph_graph_free(g)


def test_c_api_bfs_invalid_seed(emu, golden):  # noqa: F811
    c = golden["c_api"]["bfs_6"]
    g = graph(emu, c["src"], c["dst"], c["weights"])
    srcs = np.asarray([77], np.int32)
    res, err = C.c_void_p(), C.c_void_p()
    code = emu.cugraph_bfs(C.c_void_p(emu.handle), g, _view(emu, srcs), 0, C.c_size_t(10), 1, 0, C.byref(res), C.byref(err))
    assert code == 4                                                                   # CUGRAPH_INVALID_INPUT
    emu.cugraph_graph_free(g)


@pytest.mark.parametrize("wdtype", [np.float32, np.float64])
@pytest.mark.parametrize("store_transposed", [False, True])
def test_c_api_sssp_golden(emu, golden, store_transposed, wdtype):  # noqa: F811
    c = golden["c_api"]["sssp_6"]
    g = graph(emu, c["src"], c["dst"], c["weights"], store_transposed, False, wdtype)
    emu.cugraph_sssp.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    res, err = C.c_void_p(), C.c_void_p()
    code = emu.cugraph_sssp(C.c_void_p(emu.handle), g, c["source"], c["cutoff"], 1, 0, C.byref(res), C.byref(err))
    assert code == 0, emu.cugraph_error_message(err)
    verts, dist, pred = _paths(emu, res)
    big = FLT_MAX if wdtype == np.float32 else float(np.finfo(np.float64).max)
    exp = [big if x == "MAX" else x for x in c["distances"]]
    for a, b in zip(by_vertex(verts, dist, 6), exp):
        assert abs(a - b) <= 1e-3 * max(abs(a), abs(b))
    assert by_vertex(verts, pred, 6).tolist() == c["predecessors"]
    emu.cugraph_graph_free(g)


@pytest.mark.parametrize("name", ["karate.csv", "dolphins.csv", "Simple_1", "Simple_2"])
def test_pylibcugraph_sssp_goldens(emu, golden, name):  # noqa: F811
    d = golden["pylibcugraph"][name]
    s = d["sssp"]
    g = graph(emu, d["src"], d["dst"], d["weights"], False, False)
    emu.cugraph_sssp.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    res, err = C.c_void_p(), C.c_void_p()
    code = emu.cugraph_sssp(C.c_void_p(emu.handle), g, s["source"], float(s["cutoff"]), 1, 0, C.byref(res), C.byref(err))
    assert code == 0, emu.cugraph_error_message(err)
    verts, dist, pred = _paths(emu, res)
    assert verts.dtype == np.int32 and dist.dtype == np.float32 and pred.dtype == np.int32
    for v, a, p in zip(verts.tolist(), dist.tolist(), pred.tolist()):
        e = s["distances"][v]
        if a <= 3.4e38 or e <= 3.4e38:
            assert a == pytest.approx(e, 1e-4)
        if s["predecessors_checked"]:
            assert p == s["predecessors"][v]
    emu.cugraph_graph_free(g)
