"""The edge cases of tests/test_edge_cases_gpu.py and a few of tests/test_traversal_gpu.py, through the EMULATED C ABI on the
CPU: empty / tiny graphs, isolated vertices given only through the vertex list, self-loops and multi-edges kept, zero-weight
edges, BFS depth limit / several sources / isolated source, SSSP cutoff, a hub spanning many advance tiles, int64 ids,
CSR input."""
import ctypes as C

import numpy as np
import pytest

import oracle
from tests.test_emu_algorithms_cpu import _paths, run_pagerank
from tests.test_emu_staging_cpu import FLOAT32, FLOAT64, INT32, INT64, Props, emu  # noqa: F401

FLT_MAX = float(np.finfo(np.float32).max)
I32_MAX = 2**31 - 1


def _view(L, a):
    t = {np.dtype(np.int32): INT32, np.dtype(np.int64): INT64, np.dtype(np.float32): FLOAT32, np.dtype(np.float64): FLOAT64}[a.dtype]
    return C.c_void_p(L.cugraph_type_erased_device_array_view_create(a.ctypes.data, a.size, t))


class G:
    """graph + the host arrays its views point at"""

    def __init__(self, L, src, dst, w=None, vertices=None, symmetric=False, store_transposed=False, renumber=True,
                 idt=np.int32, **flags):
        self.L = L
        self.a = [None if x is None else np.ascontiguousarray(x, dt) for x, dt in
                  ((vertices, idt), (src, idt), (dst, idt), (w, np.float32 if w is None or np.asarray(w).dtype != np.float64 else np.float64))]
        v = [None if x is None else _view(L, x) for x in self.a]
        g, err = C.c_void_p(), C.c_void_p()
        code = L.cugraph_graph_create_with_times_sg(C.c_void_p(L.handle), C.byref(Props(int(symmetric), 1)), v[0], v[1], v[2], v[3],
                                                    None, None, None, None, int(store_transposed), int(renumber),
                                                    int(flags.get("drop_self_loops", 0)), int(flags.get("drop_multi_edges", 0)),
                                                    int(flags.get("symmetrize", 0)), 0, C.byref(g), C.byref(err))
        assert code == 0, L.cugraph_error_message(err)
        self.g = g

    def bfs(self, sources, direction_optimizing=False, depth_limit=0, predecessors=True, idt=np.int32):
        L = self.L
        s = np.asarray(sources, idt)
        res, err = C.c_void_p(), C.c_void_p()
        dl = depth_limit if depth_limit > 0 else I32_MAX - 1
        code = L.cugraph_bfs(C.c_void_p(L.handle), self.g, _view(L, s), int(direction_optimizing), C.c_size_t(dl), int(predecessors),
                             0, C.byref(res), C.byref(err))
        assert code == 0, L.cugraph_error_message(err)
        return _paths(L, res)

    def sssp(self, source, cutoff=float("inf")):
        L = self.L
        L.cugraph_sssp.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        res, err = C.c_void_p(), C.c_void_p()
        code = L.cugraph_sssp(C.c_void_p(L.handle), self.g, int(source), float(cutoff), 1, 0, C.byref(res), C.byref(err))
        assert code == 0, L.cugraph_error_message(err)
        return _paths(L, res)


def by_vertex(verts, vals, n):
    out = np.zeros(n, dtype=vals.dtype)
    out[verts] = vals
    return out


def test_empty_edge_list_with_vertices(emu):  # noqa: F811
    V = 5
    g = G(emu, np.zeros(0, np.int32), np.zeros(0, np.int32), vertices=np.arange(V), store_transposed=True)
    verts, vals, it = run_pagerank(emu, g.g, 0.85, 0.0, 20)
    np.testing.assert_allclose(by_vertex(verts, vals, V), np.full(V, 1.0 / V), rtol=1e-6)


def test_single_self_loop(emu):  # noqa: F811
    g = G(emu, [0], [0], store_transposed=True)
    verts, vals, it = run_pagerank(emu, g.g, 0.85, 0.0, 20)
    assert vals.size == 1 and abs(float(vals[0]) - 1.0) < 1e-6


def test_isolated_vertices_self_loops_multi_edges(emu):  # noqa: F811
    V = 40
    rng = np.random.default_rng(0)
    s = rng.integers(0, 30, 400).astype(np.int32)
    d = rng.integers(0, 30, 400).astype(np.int32)
    s[:20] = d[:20]
    s[20:60] = s[60:100]
    d[20:60] = d[60:100]
    g = G(emu, s, d, vertices=np.arange(V), store_transposed=True)
    verts, vals, _ = run_pagerank(emu, g.g, 0.85, 0.0, 30)
    ref, _, _ = oracle.pagerank(s, d, V, None, alpha=0.85, epsilon=0.0, max_iterations=30)
    np.testing.assert_allclose(by_vertex(verts, vals, V), ref, rtol=1e-5)
    bv, dist, pred = g.bfs([int(s[100])])                      # push view built lazily from the stored CSC
    rd, _ = oracle.bfs(s, d, V, [int(s[100])])
    assert np.array_equal(by_vertex(bv, dist, V), rd)


def test_sssp_zero_weight_edges_and_unreachable(emu):  # noqa: F811
    g = G(emu, [0, 1, 2, 5], [1, 2, 3, 6], np.array([0.0, 0.0, 1.5, 2.0], np.float32), vertices=np.arange(7))
    verts, dist, pred = g.sssp(0)
    assert by_vertex(verts, dist, 7).tolist() == [0.0, 0.0, 0.0, 1.5, FLT_MAX, FLT_MAX, FLT_MAX]
    p = by_vertex(verts, pred, 7)
    assert p[0] == -1 and p[4] == -1 and p[6] == -1 and p[3] == 2


def test_bfs_from_isolated_vertex_and_all_sources(emu):  # noqa: F811
    g = G(emu, [0, 1], [1, 2], vertices=np.arange(5))
    verts, dist, pred = g.bfs([4])
    got = by_vertex(verts, dist, 5)
    assert got[4] == 0 and (got[:4] == I32_MAX).all()
    verts, dist, pred = g.bfs(np.arange(5))
    assert (by_vertex(verts, dist, 5) == 0).all() and (by_vertex(verts, pred, 5) == -1).all()


def test_large_hub_spans_many_tiles(emu):  # noqa: F811
    n = 20000
    src = np.concatenate([np.zeros(n, np.int32), np.arange(1, n + 1, dtype=np.int32)])
    dst = np.concatenate([np.arange(1, n + 1, dtype=np.int32), np.zeros(n, np.int32)])
    w = np.linspace(0.5, 1.5, 2 * n).astype(np.float32)
    g = G(emu, src, dst, w, symmetric=True)
    for do in (False, True):
        verts, dist, pred = g.bfs([5], direction_optimizing=do)
        got = by_vertex(verts, dist, n + 1)
        assert got[5] == 0 and got[0] == 1 and (np.delete(got, [0, 5]) == 2).all()
    verts, dist, pred = g.sssp(0)
    rd, _ = oracle.sssp(src, dst, w, n + 1, 0)
    assert np.array_equal(by_vertex(verts, dist, n + 1).astype(np.float64), rd)


def test_bfs_depth_limit_and_no_predecessors(emu):  # noqa: F811
    g = G(emu, [0, 1, 2, 3], [1, 2, 3, 4], renumber=False)
    verts, dist, pred = g.bfs([0], depth_limit=2, predecessors=False)
    assert by_vertex(verts, dist, 5).tolist() == [0, 1, 2, I32_MAX, I32_MAX]


@pytest.mark.parametrize("do", [False, True])
def test_bfs_random_multi_source_vs_oracle(emu, do):  # noqa: F811
    rng = np.random.default_rng(9)
    V, E = 2000, 40000
    s = rng.integers(0, V, E).astype(np.int32)
    d = rng.integers(0, V, E).astype(np.int32)
    s2, d2 = np.concatenate([s, d]), np.concatenate([d, s])
    g = G(emu, s2, d2, vertices=np.arange(V), symmetric=True)
    srcs = [3, 77, 1500]
    verts, dist, pred = g.bfs(srcs, direction_optimizing=do)
    rd, _ = oracle.bfs(s2, d2, V, srcs)
    assert np.array_equal(by_vertex(verts, dist, V), rd)
    assert oracle.check_bfs_predecessors(s2, d2, V, by_vertex(verts, dist, V), by_vertex(verts, pred, V), srcs)


def test_bfs_int64_vertices(emu):  # noqa: F811
    rng = np.random.default_rng(4)
    V, E = 500, 3000
    s = rng.integers(0, V, E)
    d = rng.integers(0, V, E)
    big = 10_000_000_000
    g = G(emu, s * 3 + big, d * 3 + big, idt=np.int64)
    verts, dist, pred = g.bfs([int(s[0]) * 3 + big], idt=np.int64)
    assert verts.dtype == np.int64 and dist.dtype == np.int64
    ids, inv = np.unique(np.concatenate([s, d]), return_inverse=True)
    rd, _ = oracle.bfs(inv[:E].astype(np.int32), inv[E:].astype(np.int32), ids.size, [int(np.searchsorted(ids, s[0]))])
    got = np.zeros(ids.size, np.int64)
    got[np.searchsorted(ids * 3 + big, verts)] = dist
    reach = rd < np.iinfo(rd.dtype).max
    assert (got[reach] == rd[reach]).all() and (got[~reach] == np.iinfo(np.int64).max).all()


def test_sssp_cutoff(emu):  # noqa: F811
    rng = np.random.default_rng(6)
    V, E = 400, 3000
    s = rng.integers(0, V, E).astype(np.int32)
    d = rng.integers(0, V, E).astype(np.int32)
    w = rng.random(E).astype(np.float32)
    g = G(emu, s, d, w, vertices=np.arange(V), renumber=False)
    verts, dist, pred = g.sssp(0, cutoff=0.7)
    rd, _ = oracle.sssp(s, d, w, V, 0, cutoff=0.7)
    assert np.array_equal(by_vertex(verts, dist, V).astype(np.float64), rd)


def test_symmetrize_and_drop_flags_match_oracle_bfs(emu):  # noqa: F811
    rng = np.random.default_rng(12)
    V, E = 300, 1500
    s = rng.integers(0, V, E).astype(np.int32)
    d = rng.integers(0, V, E).astype(np.int32)
    g = G(emu, s, d, vertices=np.arange(V), symmetric=True, renumber=True, drop_self_loops=1, drop_multi_edges=1, symmetrize=1)
    verts, dist, pred = g.bfs([0])
    keep = s != d
    s2, d2 = np.concatenate([s[keep], d[keep]]), np.concatenate([d[keep], s[keep]])
    rd, _ = oracle.bfs(s2, d2, V, [0])
    assert np.array_equal(by_vertex(verts, dist, V), rd)


def test_multi_edges_are_removed_before_symmetrize(emu):  # noqa: F811
    """the reference's order (c_api/graph_sg.cpp:203-247): remove_multi_edges (minimum weight: the graph is declared
    symmetric) THEN symmetrize (reciprocal pairs are averaged): (0,1,1), (0,1,2), (1,0,5) -> (0,1,1), (1,0,5) -> weight 3;
    and symmetrize without the symmetric property is rejected (graph_sg.cpp:737-742)"""
    s = np.array([0, 0, 1, 1], np.int32)
    d = np.array([1, 1, 0, 2], np.int32)
    w = np.array([1.0, 2.0, 5.0, 4.0], np.float32)
    g = G(emu, s, d, w, vertices=np.arange(3), symmetric=True, renumber=True, drop_multi_edges=1, symmetrize=1)
    verts, dist, _ = g.sssp(2)
    assert by_vertex(verts, dist, 3).tolist() == [7.0, 4.0, 0.0]       # 2 -(4)- 1 -(3)- 0
    code, msg = _create_checked(emu, s, d, symmetric=False, multigraph=True, symmetrize=1)
    assert code != 0 and "must be symmetric if 'symmetrize'" in msg


def test_csr_input(emu):  # noqa: F811
    rng = np.random.default_rng(3)
    V, E = 200, 1500
    s = np.sort(rng.integers(0, V, E)).astype(np.int32)
    d = rng.integers(0, V, E).astype(np.int32)
    w = rng.random(E).astype(np.float32)
    off = np.zeros(V + 1, np.int32)
    off[1:] = np.cumsum(np.bincount(s, minlength=V))
    L = emu
    g, err = C.c_void_p(), C.c_void_p()
    code = L.cugraph_graph_create_sg_from_csr(C.c_void_p(L.handle), C.byref(Props(0, 1)), _view(L, off), _view(L, d), _view(L, w),
                                              None, None, 0, 0, 0, 0, C.byref(g), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    verts, pr, it = run_pagerank(L, g, 0.85, 0.0, 15)
    ref, _, _ = oracle.pagerank(s, d, V, w.astype(np.float64), alpha=0.85, epsilon=0.0, max_iterations=15)
    np.testing.assert_allclose(by_vertex(verts, pr, V), ref, rtol=2e-5)
    L.cugraph_graph_free(g)


def _create_checked(L, src, dst, symmetric, multigraph, transposed=False, renumber=True, **flags):
    """cugraph_graph_create_with_times_sg with do_expensive_check = TRUE; returns (code, message)"""
    a = [np.ascontiguousarray(src, np.int32), np.ascontiguousarray(dst, np.int32)]
    v = [_view(L, x) for x in a]
    g, err = C.c_void_p(), C.c_void_p()
    code = L.cugraph_graph_create_with_times_sg(C.c_void_p(L.handle), C.byref(Props(int(symmetric), int(multigraph))), None, v[0], v[1],
                                                None, None, None, None, None, int(transposed), int(renumber),
                                                0, int(flags.get("drop_multi_edges", 0)), int(flags.get("symmetrize", 0)), 1,
                                                C.byref(g), C.byref(err))
    msg = L.cugraph_error_message(err).decode() if code != 0 else ""
    if code == 0:
        L.cugraph_graph_free(g)
    else:
        assert not g.value
    return code, msg


@pytest.mark.parametrize("transposed", [False, True])
@pytest.mark.parametrize("renumber", [False, True])
def test_expensive_check_at_graph_creation(emu, transposed, renumber):  # noqa: F811
    """do_expensive_check of the constructors (create_graph_from_edgelist_impl.cuh:803-830); the first case is the
    reference's test_create_sg_graph_symmetric_error (cpp/tests/c_api/create_graph_test.c:430-535)"""
    src, dst = [0, 1, 1, 2, 2, 2, 3, 4], [1, 3, 4, 0, 1, 3, 5, 5]
    kw = dict(transposed=transposed, renumber=renumber)
    code, msg = _create_checked(emu, src, dst, symmetric=True, multigraph=False, **kw)
    assert code == 1 and "not symmetric" in msg                      # CUGRAPH_UNKNOWN_ERROR, as the reference's C layer reports it
    assert _create_checked(emu, src, dst, symmetric=False, multigraph=False, **kw)[0] == 0
    assert _create_checked(emu, src + dst, dst + src, symmetric=True, multigraph=False, **kw)[0] == 0
    assert _create_checked(emu, src, dst, symmetric=True, multigraph=False, symmetrize=1, **kw)[0] == 0
    code, msg = _create_checked(emu, src + [2], dst + [3], symmetric=False, multigraph=False, **kw)
    assert code == 1 and "parallel edges" in msg
    assert _create_checked(emu, src + [2], dst + [3], symmetric=False, multigraph=True, **kw)[0] == 0
    assert _create_checked(emu, src + [2], dst + [3], symmetric=False, multigraph=False, drop_multi_edges=1, **kw)[0] == 0
    # a larger random case: symmetric with parallel edges declared as a multigraph passes, one missing reverse edge fails
    r = np.random.default_rng(4)
    s, d = r.integers(0, 300, 4000), r.integers(0, 300, 4000)
    s2, d2 = np.concatenate([s, d]), np.concatenate([d, s])
    assert _create_checked(emu, s2, d2, symmetric=True, multigraph=True, **kw)[0] == 0
    keep = np.ones(s2.size, bool)
    victim = int(np.nonzero(s2 != d2)[0][0])
    keep[(s2 == s2[victim]) & (d2 == d2[victim])] = False                 # every copy of one direction
    has_reverse = ((s2 == d2[victim]) & (d2 == s2[victim]) & keep).any()
    code, msg = _create_checked(emu, s2[keep], d2[keep], symmetric=True, multigraph=True, **kw)
    assert (code == 1 and "not symmetric" in msg) if has_reverse else code == 0
