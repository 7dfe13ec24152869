"""Edge cases the reference's tests exercise implicitly: empty / tiny graphs, isolated vertices given only
through the vertex list, self-loops and multi-edges kept, zero-weight edges, repeated calls on one graph."""
import numpy as np
import pytest

import oracle
from tests.gpu_util import by_vertex, make_graph

pytestmark = pytest.mark.gpu


def _pr(h, g, iters=20, alpha=0.85):
    from cugraph_b200 import pylibcugraph as plc
    return plc.pagerank(h, g, None, None, None, None, alpha, 0.0, iters, False, fail_on_nonconvergence=False)


def test_empty_edge_list_with_vertices():
    V = 5
    h, g = make_graph([], [], vertices=np.arange(V, dtype=np.int32), store_transposed=True)
    verts, vals, conv = _pr(h, g)
    got = by_vertex(verts, vals, V)
    np.testing.assert_allclose(got, np.full(V, 1.0 / V), rtol=1e-6)   # every vertex dangling: stays uniform


def test_single_self_loop():
    h, g = make_graph([0], [0], store_transposed=True)
    verts, vals, conv = _pr(h, g)
    assert vals.numel() == 1 and abs(float(vals[0]) - 1.0) < 1e-6


def test_isolated_vertices_self_loops_multi_edges():
    V = 40
    rng = np.random.default_rng(0)
    s = rng.integers(0, 30, 400).astype(np.int32)   # vertices 30..39 isolated
    d = rng.integers(0, 30, 400).astype(np.int32)
    s[:20] = d[:20]                                  # self loops
    s[20:60] = s[60:100]
    d[20:60] = d[60:100]                             # multi-edges
    h, g = make_graph(s, d, vertices=np.arange(V, dtype=np.int32), store_transposed=True)
    verts, vals, _ = _pr(h, g, 30)
    ref, _, _ = oracle.pagerank(s, d, V, None, alpha=0.85, epsilon=0.0, max_iterations=30)
    np.testing.assert_allclose(by_vertex(verts, vals, V), ref, rtol=1e-6)
    # the same graph object serves BFS (push view is built lazily from the stored CSC)
    import torch
    from cugraph_b200 import pylibcugraph as plc
    dist, pred, bv = plc.bfs(h, g, torch.tensor([int(s[100])], dtype=torch.int32).cuda(), False, 0, True, False)
    rd, _ = oracle.bfs(s, d, V, [int(s[100])])
    assert np.array_equal(by_vertex(bv, dist, V), rd)


def test_repeated_calls_are_deterministic_enough():
    rng = np.random.default_rng(2)
    V, E = 3000, 60000
    s = rng.integers(0, V, E).astype(np.int32)
    d = rng.integers(0, V, E).astype(np.int32)
    h, g = make_graph(s, d, vertices=np.arange(V, dtype=np.int32), store_transposed=True)
    a = by_vertex(*_pr(h, g, 25)[:2], V)
    b = by_vertex(*_pr(h, g, 25)[:2], V)
    np.testing.assert_allclose(a, b, rtol=2e-7)   # fp64 atomics may reorder; rounding to fp32 hides it


def test_sssp_zero_weight_edges_and_unreachable():
    src = [0, 1, 2, 5]
    dst = [1, 2, 3, 6]
    w = [0.0, 0.0, 1.5, 2.0]
    h, g = make_graph(src, dst, w, vertices=np.arange(7, dtype=np.int32))
    from cugraph_b200 import pylibcugraph as plc
    verts, dist, pred = plc.sssp(h, g, 0, float("inf"), True, False)
    got = by_vertex(verts, dist, 7)
    fmax = float(np.finfo(np.float32).max)
    assert got.tolist() == [0.0, 0.0, 0.0, 1.5, fmax, fmax, fmax]
    p = by_vertex(verts, pred, 7)
    assert p[0] == -1 and p[4] == -1 and p[6] == -1 and p[3] == 2


def test_bfs_from_isolated_vertex_and_all_sources():
    import torch
    from cugraph_b200 import pylibcugraph as plc
    src = [0, 1]
    dst = [1, 2]
    h, g = make_graph(src, dst, vertices=np.arange(5, dtype=np.int32))
    dist, pred, verts = plc.bfs(h, g, torch.tensor([4], dtype=torch.int32).cuda(), False, 0, True, False)
    got = by_vertex(verts, dist, 5)
    assert got[4] == 0 and (got[:4] == 2**31 - 1).all()
    dist, pred, verts = plc.bfs(h, g, torch.arange(5, dtype=torch.int32).cuda(), False, 0, True, False)
    assert (by_vertex(verts, dist, 5) == 0).all() and (by_vertex(verts, pred, 5) == -1).all()


def test_large_hub_goes_through_large_queue():
    """A star with 20000 leaves: the hub exceeds the large-degree threshold of the advance."""
    n = 20000
    src = np.concatenate([np.zeros(n, dtype=np.int32), np.arange(1, n + 1, dtype=np.int32)])
    dst = np.concatenate([np.arange(1, n + 1, dtype=np.int32), np.zeros(n, dtype=np.int32)])
    w = np.linspace(0.5, 1.5, 2 * n).astype(np.float32)
    h, g = make_graph(src, dst, w, symmetric=True)
    import torch
    from cugraph_b200 import pylibcugraph as plc
    for do in (False, True):
        dist, pred, verts = plc.bfs(h, g, torch.tensor([5], dtype=torch.int32).cuda(), do, 0, True, False)
        got = by_vertex(verts, dist, n + 1)
        assert got[5] == 0 and got[0] == 1 and (np.delete(got, [0, 5]) == 2).all()
    verts, dist, pred = plc.sssp(h, g, 0, float("inf"), True, False)
    rd, _ = oracle.sssp(src, dst, w, n + 1, 0)
    assert np.array_equal(by_vertex(verts, dist, n + 1).astype(np.float64), rd)


@pytest.mark.parametrize("store_transposed", [False, True])
def test_expensive_check_at_graph_creation(store_transposed):
    """do_expensive_check of the constructors (create_graph_from_edgelist_impl.cuh:803-830); the first case is the
    reference's test_create_sg_graph_symmetric_error (cpp/tests/c_api/create_graph_test.c:430-535)"""
    import torch
    from cugraph_b200 import _capi
    from cugraph_b200 import pylibcugraph as plc
    h = plc.ResourceHandle()
    src, dst = [0, 1, 1, 2, 2, 2, 3, 4], [1, 3, 4, 0, 1, 3, 5, 5]

    def create(s, d, symmetric, multigraph, **kw):
        return plc.SGGraph(h, plc.GraphProperties(is_symmetric=symmetric, is_multigraph=multigraph),
                           torch.tensor(s, dtype=torch.int32).cuda(), torch.tensor(d, dtype=torch.int32).cuda(),
                           store_transposed=store_transposed, renumber=True, do_expensive_check=True, **kw)

    with pytest.raises(_capi.CugraphError) as e:
        create(src, dst, True, False)
    assert e.value.code == _capi.UNKNOWN_ERROR and "not symmetric" in str(e.value)
    create(src, dst, False, False)
    create(src + dst, dst + src, True, False)
    create(src, dst, True, False, symmetrize=True)
    with pytest.raises(_capi.CugraphError) as e:
        create(src + [2], dst + [3], False, False)
    assert "parallel edges" in str(e.value)
    create(src + [2], dst + [3], False, True)
    create(src + [2], dst + [3], False, False, drop_multi_edges=True)
    # RMAT-14 symmetrised (multi-edges present): passes as a multigraph, fails once one direction of an edge is removed
    from oracle.rmat import rmat_edgelist
    s, d = rmat_edgelist(14, 16 << 14, seed=5)
    s2, d2 = np.concatenate([s, d]), np.concatenate([d, s])
    create(s2.tolist(), d2.tolist(), True, True)
    victim = int(np.nonzero(s2 != d2)[0][0])
    keep = ~((s2 == s2[victim]) & (d2 == d2[victim]))
    with pytest.raises(_capi.CugraphError):
        create(s2[keep].tolist(), d2[keep].tolist(), True, True)
