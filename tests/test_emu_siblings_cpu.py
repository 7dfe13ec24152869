"""Katz centrality, HITS and weakly connected components (the sibling algorithms on the same primitive)
through the emulated C ABI against the numpy restatements of the reference tests' CPU references (oracle.katz / .hits /
.wcc).  The reference's own katz_test.c / hits_test.c / weakly_connected_components_test.c run unmodified against the library in
tests/test_reference_c_tests_{cpu,gpu}.py."""
import ctypes as C

import numpy as np
import pytest

import oracle
from tests.test_emu_algorithms_cpu import _view_to_np, create_sym_graph, dense_ids, symmetric_edges
from tests.test_emu_staging_cpu import FLOAT32, INT32, create_graph, emu, make_edges  # noqa: F401


def _centrality(L, res):
    for f in ("cugraph_centrality_result_get_vertices", "cugraph_centrality_result_get_values"):
        getattr(L, f).restype = C.c_void_p
        getattr(L, f).argtypes = [C.c_void_p]
    L.cugraph_centrality_result_free.argtypes = [C.c_void_p]
    v = _view_to_np(L, L.cugraph_centrality_result_get_vertices(res))
    x = _view_to_np(L, L.cugraph_centrality_result_get_values(res))
    L.cugraph_centrality_result_free(res)
    return v, x


@pytest.mark.parametrize("weighted,min_edges", [(False, "0"), (True, "1000000000")])
def test_katz_emulated(emu, monkeypatch, weighted, min_edges):  # noqa: F811
    monkeypatch.setenv("CUGRAPH_B200_SWEEP_MIN_EDGES", min_edges)   # 0: the piece-stream sweep, else the plain one
    L = emu
    L.emu_reload_tuning(C.c_void_p(L.handle))
    src, dst, w = make_edges(3_000, 40_000, seed=17, weighted=weighted)
    g = create_graph(L, src, dst, w)
    ids, inv = np.unique(np.concatenate([src, dst]), return_inverse=True)
    s, d = inv[:src.size], inv[src.size:]
    deg_max = np.bincount(d).max()
    alpha = 0.5 / (deg_max * (float(w.max()) if weighted else 1.0))
    L.cugraph_katz_centrality.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_size_t, C.c_int,
                                          C.c_void_p, C.c_void_p]
    res, err = C.c_void_p(), C.c_void_p()
    code = L.cugraph_katz_centrality(C.c_void_p(L.handle), g, None, alpha, 1.0, 1e-4, 200, 0, C.byref(res), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    verts, x = _centrality(L, res)
    ref, _ = oracle.katz(s, d, ids.size, w if weighted else None, alpha=alpha, beta=1.0, epsilon=1e-4, dtype=np.float32)
    got = np.zeros(ids.size)
    got[np.searchsorted(ids, verts)] = x
    np.testing.assert_allclose(got, ref, rtol=2e-5)
    # the iteration budget is enforced with the reference's message
    code = L.cugraph_katz_centrality(C.c_void_p(L.handle), g, None, alpha, 1.0, 0.0, 2, 0, C.byref(res), C.byref(err))
    assert code != 0 and b"failed to converge" in L.cugraph_error_message(err)
    L.cugraph_graph_free(g)
    monkeypatch.delenv("CUGRAPH_B200_SWEEP_MIN_EDGES")
    L.emu_reload_tuning(C.c_void_p(L.handle))


@pytest.mark.parametrize("transposed,weighted,normalize,guess", [(False, False, True, False), (True, True, False, True)])
def test_hits_emulated(emu, transposed, weighted, normalize, guess):  # noqa: F811
    L = emu
    src, dst, w = make_edges(2_500, 30_000, seed=23, weighted=weighted)
    g = create_graph(L, src, dst, w, store_transposed=int(transposed))
    ids, inv = np.unique(np.concatenate([src, dst]), return_inverse=True)
    s, d = inv[:src.size], inv[src.size:]
    L.cugraph_hits.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                               C.c_void_p]
    for f in ("vertices", "hubs", "authorities"):
        getattr(L, f"cugraph_hits_result_get_{f}").restype = C.c_void_p
        getattr(L, f"cugraph_hits_result_get_{f}").argtypes = [C.c_void_p]
    L.cugraph_hits_result_get_number_of_iterations.restype = C.c_size_t
    L.cugraph_hits_result_get_number_of_iterations.argtypes = [C.c_void_p]
    L.cugraph_hits_result_free.argtypes = [C.c_void_p]
    gv = gx = None
    init = None
    keep = []
    if guess:
        r = np.random.default_rng(3)
        gvert = ids.astype(np.int32)
        gval = (r.random(ids.size) + 0.1).astype(np.float32)
        keep = [gvert, gval]
        gv = C.c_void_p(L.cugraph_type_erased_device_array_view_create(gvert.ctypes.data, gvert.size, INT32))
        gx = C.c_void_p(L.cugraph_type_erased_device_array_view_create(gval.ctypes.data, gval.size, FLOAT32))
        init = gval.astype(np.float64)
    res, err = C.c_void_p(), C.c_void_p()
    code = L.cugraph_hits(C.c_void_p(L.handle), g, 1e-6, 500, gv, gx, int(normalize), 1, C.byref(res), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    verts = _view_to_np(L, L.cugraph_hits_result_get_vertices(res))
    hubs = _view_to_np(L, L.cugraph_hits_result_get_hubs(res))
    auth = _view_to_np(L, L.cugraph_hits_result_get_authorities(res))
    iters = L.cugraph_hits_result_get_number_of_iterations(res)
    L.cugraph_hits_result_free(res)
    rh, ra, rit, _ = oracle.hits(s, d, ids.size, epsilon=1e-6, initial_hubs=init, normalize=normalize)
    gh, ga = np.zeros(ids.size), np.zeros(ids.size)
    gh[np.searchsorted(ids, verts)] = hubs
    ga[np.searchsorted(ids, verts)] = auth
    assert abs(int(iters) - rit) <= 2                      # fp32 sums vs fp64: the stopping iteration may differ by a step
    np.testing.assert_allclose(gh, rh, rtol=2e-3, atol=1e-7)
    np.testing.assert_allclose(ga, ra, rtol=2e-3, atol=1e-7)
    L.cugraph_graph_free(g)
    del keep


def test_wcc_emulated(emu):  # noqa: F811
    L = emu
    s, d = symmetric_edges(6_000, 5_000, seed=31)                  # sparse: hundreds of components
    g = create_sym_graph(L, s, d, None)
    ids, ss, dd = dense_ids(s, d)
    L.cugraph_weakly_connected_components.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    for f in ("cugraph_labeling_result_get_vertices", "cugraph_labeling_result_get_labels"):
        getattr(L, f).restype = C.c_void_p
        getattr(L, f).argtypes = [C.c_void_p]
    L.cugraph_labeling_result_free.argtypes = [C.c_void_p]
    res, err = C.c_void_p(), C.c_void_p()
    code = L.cugraph_weakly_connected_components(C.c_void_p(L.handle), g, 0, C.byref(res), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    verts = _view_to_np(L, L.cugraph_labeling_result_get_vertices(res))
    labels = _view_to_np(L, L.cugraph_labeling_result_get_labels(res))
    L.cugraph_labeling_result_free(res)
    ref = oracle.wcc(ss, dd, ids.size)
    got = np.zeros(ids.size, dtype=np.int64)
    got[np.searchsorted(ids, verts)] = labels
    assert len(set(ref.tolist())) > 50
    # same partition: the map reference component -> label is a bijection, and every label is a vertex of its component
    pairs = set(zip(ref.tolist(), got.tolist()))
    assert len(pairs) == len(set(ref.tolist())) == len(set(got.tolist()))
    assert all(ref[np.searchsorted(ids, lab)] == comp for comp, lab in pairs)
    L.cugraph_graph_free(g)
    # a directed graph is rejected with the reference's message
    src, dst, _ = make_edges(500, 2_000, seed=2)
    g2 = create_graph(L, src, dst, None)
    code = L.cugraph_weakly_connected_components(C.c_void_p(L.handle), g2, 0, C.byref(res), C.byref(err))
    assert code != 0 and b"should be symmetric" in L.cugraph_error_message(err)
    L.cugraph_graph_free(g2)
