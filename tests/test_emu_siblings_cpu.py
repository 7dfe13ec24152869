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


@pytest.mark.parametrize("weighted", [False, True])
def test_eigenvector_centrality_emulated(emu, weighted):  # noqa: F811
    L = emu
    src, dst, w = make_edges(2_000, 30_000, seed=29, weighted=weighted)
    g = create_graph(L, src, dst, w)
    ids, inv = np.unique(np.concatenate([src, dst]), return_inverse=True)
    s, d = inv[:src.size], inv[src.size:]
    L.cugraph_eigenvector_centrality.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
    res, err = C.c_void_p(), C.c_void_p()
    code = L.cugraph_eigenvector_centrality(C.c_void_p(L.handle), g, 1e-7, 1000, 0, C.byref(res), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    verts, x = _centrality(L, res)
    ref, _ = oracle.eigenvector(s, d, ids.size, w if weighted else None, epsilon=1e-7, max_iterations=1000)
    got = np.zeros(ids.size)
    got[np.searchsorted(ids, verts)] = x
    np.testing.assert_allclose(got, ref, rtol=2e-3, atol=1e-7)
    code = L.cugraph_eigenvector_centrality(C.c_void_p(L.handle), g, 0.0, 3, 0, C.byref(res), C.byref(err))
    assert code != 0 and b"failed to converge" in L.cugraph_error_message(err)
    L.cugraph_graph_free(g)


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


def test_int64_ids_double_weights_and_empty_graph(emu):  # noqa: F811
    """64-bit external ids, fp64 weights (results come back in the graph's types), renumber = FALSE reporting order, and an
    edgeless graph, through Katz / HITS / components / extract_paths"""
    from tests.test_emu_edge_cases_cpu import G, _view
    L = emu
    r = np.random.default_rng(8)
    V, E = 300, 2500
    big = 5_000_000_000
    a, b = r.integers(0, V, E), r.integers(0, V, E)
    s = np.concatenate([a, b]) * 7 + big
    d = np.concatenate([b, a]) * 7 + big
    w = np.concatenate([r.random(E) + 0.5] * 2)
    g = G(L, s, d, w, symmetric=True, store_transposed=False, idt=np.int64)
    ids, inv = np.unique(np.concatenate([s, d]), return_inverse=True)
    si, di = inv[:s.size], inv[s.size:]
    # Katz in fp64
    L.cugraph_katz_centrality.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_size_t, C.c_int,
                                          C.c_void_p, C.c_void_p]
    alpha = 0.4 / (np.bincount(di).max() * float(w.max()))
    res, err = C.c_void_p(), C.c_void_p()
    code = L.cugraph_katz_centrality(C.c_void_p(L.handle), g.g, None, alpha, 1.0, 1e-10, 500, 0, C.byref(res), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    verts, x = _centrality(L, res)
    assert verts.dtype == np.int64 and x.dtype == np.float64
    ref, _ = oracle.katz(si, di, ids.size, w, alpha=alpha, beta=1.0, epsilon=1e-10)
    got = np.zeros(ids.size)
    got[np.searchsorted(ids, verts)] = x
    np.testing.assert_allclose(got, ref, rtol=1e-9)
    # components with 64-bit labels
    L.cugraph_weakly_connected_components.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    for f in ("cugraph_labeling_result_get_vertices", "cugraph_labeling_result_get_labels"):
        getattr(L, f).restype = C.c_void_p
        getattr(L, f).argtypes = [C.c_void_p]
    L.cugraph_labeling_result_free.argtypes = [C.c_void_p]
    code = L.cugraph_weakly_connected_components(C.c_void_p(L.handle), g.g, 0, C.byref(res), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    lv = _view_to_np(L, L.cugraph_labeling_result_get_vertices(res))
    ll = _view_to_np(L, L.cugraph_labeling_result_get_labels(res))
    L.cugraph_labeling_result_free(res)
    assert lv.dtype == np.int64 and ll.dtype == np.int64 and set(ll.tolist()) <= set(ids.tolist())
    comp = oracle.wcc(si, di, ids.size)
    gl = np.zeros(ids.size, dtype=np.int64)
    gl[np.searchsorted(ids, lv)] = ll
    assert len(set(zip(comp.tolist(), gl.tolist()))) == len(set(comp.tolist())) == len(set(gl.tolist()))
    # BFS + extract_paths with 64-bit ids
    src_v = int(ids[np.bincount(si).argmax()])
    verts, dist, pred = None, None, None
    sv = np.array([src_v], dtype=np.int64)
    bres = C.c_void_p()
    code = L.cugraph_bfs(C.c_void_p(L.handle), g.g, _view(L, sv), 0, C.c_size_t(2**31 - 2), 1, 0, C.byref(bres), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    for f in ("cugraph_paths_result_get_vertices", "cugraph_paths_result_get_distances"):
        getattr(L, f).restype = C.c_void_p
        getattr(L, f).argtypes = [C.c_void_p]
    bv = _view_to_np(L, L.cugraph_paths_result_get_vertices(bres))
    bd = _view_to_np(L, L.cugraph_paths_result_get_distances(bres))
    dist_of = dict(zip(bv.tolist(), bd.tolist()))
    dests = ids[r.integers(0, ids.size, 40)].astype(np.int64)
    L.cugraph_extract_paths.argtypes = [C.c_void_p] * 7
    L.cugraph_extract_paths_result_get_max_path_length.restype = C.c_size_t
    L.cugraph_extract_paths_result_get_max_path_length.argtypes = [C.c_void_p]
    L.cugraph_extract_paths_result_get_paths.restype = C.c_void_p
    L.cugraph_extract_paths_result_get_paths.argtypes = [C.c_void_p]
    L.cugraph_extract_paths_result_free.argtypes = [C.c_void_p]
    out = C.c_void_p()
    code = L.cugraph_extract_paths(C.c_void_p(L.handle), g.g, _view(L, sv), bres, _view(L, dests), C.byref(out), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    length = int(L.cugraph_extract_paths_result_get_max_path_length(out))
    paths = _view_to_np(L, L.cugraph_extract_paths_result_get_paths(out)).reshape(dests.size, length)
    assert paths.dtype == np.int64
    i64max = np.iinfo(np.int64).max
    edges = set(zip(s.tolist(), d.tolist()))
    for row, t in zip(paths, dests.tolist()):
        dt = dist_of[t]
        if dt == i64max:
            assert (row == -1).all()
        else:
            assert row[0] == src_v and row[dt] == t and all((int(p), int(q)) in edges for p, q in zip(row[:dt], row[1:dt + 1]))
    L.cugraph_extract_paths_result_free(out)
    L.cugraph_paths_result_free(bres)
    L.cugraph_graph_free(g.g)
    # a graph with vertices and no edges: Katz gives beta / ||.||, components one label per vertex, HITS reports a zero norm
    g0 = G(L, np.zeros(0, np.int32), np.zeros(0, np.int32), vertices=np.arange(6), symmetric=True, store_transposed=True)
    code = L.cugraph_katz_centrality(C.c_void_p(L.handle), g0.g, None, 0.1, 1.0, 1e-6, 10, 0, C.byref(res), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    _, x0 = _centrality(L, res)
    np.testing.assert_allclose(x0, np.full(6, 1.0 / np.sqrt(6.0)), rtol=1e-6)
    code = L.cugraph_weakly_connected_components(C.c_void_p(L.handle), g0.g, 0, C.byref(res), C.byref(err))
    assert code == 0
    l0 = _view_to_np(L, L.cugraph_labeling_result_get_labels(res))
    L.cugraph_labeling_result_free(res)
    assert sorted(l0.tolist()) == list(range(6))
    L.cugraph_hits.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                               C.c_void_p]
    code = L.cugraph_hits(C.c_void_p(L.handle), g0.g, 1e-6, 10, None, None, 1, 0, C.byref(res), C.byref(err))
    assert code != 0 and b"positive" in L.cugraph_error_message(err)      # "Norm is required to be a positive value."
    L.cugraph_graph_free(g0.g)


@pytest.mark.parametrize("store_transposed", [0, 1])
def test_degrees_emulated(emu, store_transposed):  # noqa: F811
    """cugraph_degrees / _in_degrees / _out_degrees on a directed multigraph, for all vertices and for a subset"""
    L = emu
    src, dst, _ = make_edges(1_500, 12_000, seed=41)
    g = create_graph(L, src, dst, None, store_transposed=store_transposed)
    ids = np.unique(np.concatenate([src, dst]))
    indeg = dict(zip(*np.unique(dst, return_counts=True)))
    outdeg = dict(zip(*np.unique(src, return_counts=True)))
    for f in ("cugraph_in_degrees", "cugraph_out_degrees", "cugraph_degrees"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    for f in ("vertices", "in_degrees", "out_degrees"):
        getattr(L, f"cugraph_degrees_result_get_{f}").restype = C.c_void_p
        getattr(L, f"cugraph_degrees_result_get_{f}").argtypes = [C.c_void_p]
    L.cugraph_degrees_result_free.argtypes = [C.c_void_p]
    subset = np.ascontiguousarray(ids[::7][::-1], dtype=np.int32)
    sub_view = C.c_void_p(L.cugraph_type_erased_device_array_view_create(subset.ctypes.data, subset.size, INT32))
    for fn, want_in, want_out in (("cugraph_degrees", True, True), ("cugraph_in_degrees", True, False), ("cugraph_out_degrees", False, True)):
        for sv, expect_v in ((None, None), (sub_view, subset)):
            res, err = C.c_void_p(), C.c_void_p()
            code = getattr(L, fn)(C.c_void_p(L.handle), g, sv, 0, C.byref(res), C.byref(err))
            assert code == 0, L.cugraph_error_message(err)
            v = _view_to_np(L, L.cugraph_degrees_result_get_vertices(res))
            pin, pout = L.cugraph_degrees_result_get_in_degrees(res), L.cugraph_degrees_result_get_out_degrees(res)
            assert bool(pin) == want_in and bool(pout) == want_out
            if expect_v is None:
                assert sorted(v.tolist()) == ids.tolist()
            else:
                assert v.tolist() == expect_v.tolist()
            if pin:
                assert _view_to_np(L, pin).tolist() == [int(indeg.get(x, 0)) for x in v.tolist()]
            if pout:
                assert _view_to_np(L, pout).tolist() == [int(outdeg.get(x, 0)) for x in v.tolist()]
            L.cugraph_degrees_result_free(res)
    L.cugraph_graph_free(g)
