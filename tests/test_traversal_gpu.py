"""BFS / SSSP parity through the C-ABI: distances bit-exact against the CPU oracle, predecessors
checked with the validity predicates the reference's own tests use (bfs_test.cpp:213-233,
sssp_test.cpp:222-240) and bit-exact on the golden fixtures where the parent is unique."""
import numpy as np
import pytest

import oracle
from oracle.rmat import rmat_edgelist
from tests.gpu_util import by_vertex, make_graph

pytestmark = pytest.mark.gpu

INT_MAX = 2**31 - 1
FLT_MAX = float(np.finfo(np.float32).max)


def _bfs(h, g, sources, do=False, depth_limit=0, pred=True, dtype=np.int32):
    import torch
    from cugraph_b200 import pylibcugraph as plc
    s = torch.as_tensor(np.asarray(sources, dtype=dtype)).cuda()
    return plc.bfs(h, g, s, do, depth_limit, pred, False)


@pytest.mark.parametrize("store_transposed", [False, True])
@pytest.mark.parametrize("renumber", [False, True])
def test_bfs_c_api_golden(golden, store_transposed, renumber):
    c = golden["c_api"]["bfs_6"]
    h, g = make_graph(c["src"], c["dst"], c["weights"], store_transposed=store_transposed, renumber=renumber)
    dist, pred, verts = _bfs(h, g, c["sources"], depth_limit=c["depth_limit"])
    assert by_vertex(verts, dist, 6).tolist() == c["distances"]
    assert by_vertex(verts, pred, 6).tolist() == c["predecessors"]


def test_bfs_invalid_source_and_type(golden):
    from cugraph_b200 import _capi
    c = golden["c_api"]["bfs_6"]
    h, g = make_graph(c["src"], c["dst"], c["weights"])
    with pytest.raises(_capi.CugraphError) as e:  # bfs_test.c:108-157: INT64 seeds on an INT32 graph
        _bfs(h, g, [0], dtype=np.int64)
    assert e.value.code == _capi.INVALID_INPUT
    with pytest.raises(_capi.CugraphError) as e:
        _bfs(h, g, [77])
    assert e.value.code == _capi.INVALID_INPUT


def test_bfs_depth_limit_and_no_predecessors():
    src = [0, 1, 2, 3]
    dst = [1, 2, 3, 4]
    h, g = make_graph(src, dst)
    dist, pred, verts = _bfs(h, g, [0], depth_limit=2, pred=False)
    assert by_vertex(verts, dist, 5).tolist() == [0, 1, 2, INT_MAX, INT_MAX]
    assert pred.numel() == 0


def _sym(s, d):
    return np.concatenate([s, d]), np.concatenate([d, s])


@pytest.mark.parametrize("do", [False, True])
@pytest.mark.parametrize("V,E,nsrc", [(50, 80, 1), (2000, 6000, 1), (2000, 40000, 3), (30000, 200000, 1)])
def test_bfs_random_vs_oracle(V, E, nsrc, do):
    rng = np.random.default_rng(V + E)
    s = rng.integers(0, V, E).astype(np.int32)
    d = rng.integers(0, V, E).astype(np.int32)
    s, d = _sym(s, d)
    srcs = rng.choice(V, nsrc, replace=False).astype(np.int32)
    h, g = make_graph(s, d, symmetric=True, vertices=np.arange(V, dtype=np.int32))
    dist, pred, verts = _bfs(h, g, srcs, do=do)
    ref_d, _ = oracle.bfs(s, d, V, srcs)
    got_d = by_vertex(verts, dist, V)
    got_p = by_vertex(verts, pred, V)
    assert np.array_equal(got_d, ref_d)
    assert oracle.check_bfs_predecessors(s, d, V, got_d, got_p, srcs)


def test_bfs_directed_not_symmetric():
    rng = np.random.default_rng(9)
    V, E = 3000, 12000
    s = rng.integers(0, V, E).astype(np.int32)
    d = rng.integers(0, V, E).astype(np.int32)
    for st in (False, True):
        h, g = make_graph(s, d, store_transposed=st, vertices=np.arange(V, dtype=np.int32))
        dist, pred, verts = _bfs(h, g, [5])
        ref_d, _ = oracle.bfs(s, d, V, [5])
        got_d = by_vertex(verts, dist, V)
        assert np.array_equal(got_d, ref_d)
        assert oracle.check_bfs_predecessors(s, d, V, got_d, by_vertex(verts, pred, V), [5])
    from cugraph_b200 import _capi
    with pytest.raises(_capi.CugraphError):  # direction optimising needs a symmetric graph (bfs_impl.cuh:202-204)
        _bfs(h, g, [5], do=True)


@pytest.mark.parametrize("scale", [12, 16, 18])
def test_bfs_rmat_direction_optimizing(scale):
    s, d = rmat_edgelist(scale, 16 << scale, seed=scale + 100)
    s, d = _sym(s, d)
    V = 1 << scale
    h, g = make_graph(s, d, symmetric=True, vertices=np.arange(V, dtype=np.int32))
    deg = np.bincount(s, minlength=V)
    rng = np.random.default_rng(1)
    cand = np.nonzero(deg > 0)[0]
    csr = oracle.coo_to_csx(s, d, V)
    for src in rng.choice(cand, 3, replace=False):
        dist, pred, verts = _bfs(h, g, [src], do=True)
        ref_d, _ = oracle.bfs(s, d, V, [src], csr=csr)
        got_d = by_vertex(verts, dist, V)
        assert np.array_equal(got_d, ref_d)
        assert oracle.check_bfs_predecessors(s, d, V, got_d, by_vertex(verts, pred, V), [src])


def test_bfs_int64_vertices():
    rng = np.random.default_rng(21)
    V, E = 500, 3000
    ids = (np.arange(V, dtype=np.int64) * 1000003 + 7)
    s = rng.integers(0, V, E)
    d = rng.integers(0, V, E)
    h, g = make_graph(ids[s], ids[d], vertex_dtype=np.int64, vertices=ids)
    dist, pred, verts = _bfs(h, g, [ids[3]], dtype=np.int64)
    assert str(dist.dtype) == "torch.int64" and str(pred.dtype) == "torch.int64"
    ref_d, _ = oracle.bfs(s, d, V, [3])
    got = dict(zip(verts.tolist(), dist.tolist()))
    for i in range(V):
        exp = ref_d[i] if ref_d[i] != INT_MAX else 2**63 - 1
        assert got[int(ids[i])] == exp


# ------------------------------------------------------------------------------------------ SSSP
def _sssp(h, g, source, cutoff=float("inf"), pred=True):
    from cugraph_b200 import pylibcugraph as plc
    return plc.sssp(h, g, source, cutoff, pred, False)


@pytest.mark.parametrize("store_transposed", [False, True])
@pytest.mark.parametrize("wdtype", [np.float32, np.float64])
def test_sssp_c_api_golden(golden, store_transposed, wdtype):
    c = golden["c_api"]["sssp_6"]
    h, g = make_graph(c["src"], c["dst"], c["weights"], store_transposed=store_transposed, renumber=False,
                      weight_dtype=wdtype)
    verts, dist, pred = _sssp(h, g, c["source"], c["cutoff"])
    big = FLT_MAX if wdtype == np.float32 else float(np.finfo(np.float64).max)
    exp = [big if x == "MAX" else x for x in c["distances"]]
    got = by_vertex(verts, dist, 6)
    for a, b in zip(got, exp):
        assert abs(a - b) <= 1e-3 * max(abs(a), abs(b))
    assert by_vertex(verts, pred, 6).tolist() == c["predecessors"]


@pytest.mark.parametrize("name", ["karate.csv", "dolphins.csv", "Simple_1", "Simple_2"])
def test_sssp_pylibcugraph_golden(golden, name):
    d = golden["pylibcugraph"][name]
    s = d["sssp"]
    h, g = make_graph(d["src"], d["dst"], d["weights"], store_transposed=False, renumber=False)
    verts, dist, pred = _sssp(h, g, s["source"], s["cutoff"])
    assert str(verts.dtype) == "torch.int32" and str(dist.dtype) == "torch.float32" and str(pred.dtype) == "torch.int32"
    av, ad, ap = verts.tolist(), dist.tolist(), pred.tolist()
    for i in range(len(s["distances"])):
        e = s["distances"][av[i]]
        if ad[i] <= 3.4e38 or e <= 3.4e38:
            assert ad[i] == pytest.approx(e, 1e-4)
        if s["predecessors_checked"]:
            assert ap[i] == s["predecessors"][av[i]]


@pytest.mark.parametrize("use_float", [True, False])
@pytest.mark.parametrize("V,E", [(60, 150), (3000, 20000), (20000, 300000)])
def test_sssp_random_vs_oracle(V, E, use_float):
    rng = np.random.default_rng(V * 3 + E)
    s = rng.integers(0, V, E).astype(np.int32)
    d = rng.integers(0, V, E).astype(np.int32)
    wd = np.float32 if use_float else np.float64
    w = rng.random(E).astype(wd)
    h, g = make_graph(s, d, w, weight_dtype=wd, vertices=np.arange(V, dtype=np.int32))
    verts, dist, pred = _sssp(h, g, 1)
    ref_d, _ = oracle.sssp(s, d, w, V, 1, use_float=use_float)
    got_d = by_vertex(verts, dist, V).astype(np.float64)
    # the distance fixpoint of monotone fp add/min is order independent: bit-exact
    assert np.array_equal(got_d, ref_d)
    assert oracle.check_sssp_predecessors(s, d, w, V, got_d, by_vertex(verts, pred, V), 1)


@pytest.mark.parametrize("wdtype", [np.float32, np.float64])
def test_sssp_zero_weight_predecessors_form_a_tree(wdtype):
    """symmetric zero-weight edges, a zero-weight cycle and a weight absorbed by float rounding: both directions of such
    an edge are tight, the predecessors must still lead every reached vertex back to the source.  float32 records the
    predecessor at the relaxation (packed word), float64 derives it from the distance fixpoint (strict pass + tie passes)."""
    r = np.random.default_rng(3)
    V = 4000
    hs = r.integers(0, V, 16000).astype(np.int32)
    hd = r.integers(0, V, 16000).astype(np.int32)
    hw = np.where(r.random(16000) < 0.5, 0.0, r.random(16000)).astype(np.float32)
    extra = [(6, 7, 0.0), (7, 8, 0.0), (8, 9, 0.0), (9, 7, 0.0), (0, 3990, 1e8), (3990, 3991, 1.0), (3991, 3992, 1.0)]
    hs = np.concatenate([hs, np.array([e[0] for e in extra], np.int32)])
    hd = np.concatenate([hd, np.array([e[1] for e in extra], np.int32)])
    hw = np.concatenate([hw, np.array([e[2] for e in extra], np.float32)])
    if wdtype == np.float64:  # 1e16 + 1 == 1e16 in double
        hw = hw.astype(np.float64)
        hw[hw == 1e8] = 1e16
    s, d, w = np.concatenate([hs, hd]), np.concatenate([hd, hs]), np.concatenate([hw, hw])
    h, g = make_graph(s, d, w, symmetric=True, vertices=np.arange(V, dtype=np.int32), weight_dtype=wdtype)
    for source in (0, 7):
        verts, dist, pred = _sssp(h, g, source)
        ref_d, _ = oracle.sssp(s, d, w, V, source, use_float=(wdtype == np.float32))
        got_d, got_p = by_vertex(verts, dist, V), by_vertex(verts, pred, V)
        assert np.array_equal(got_d.astype(np.float64), ref_d)
        assert oracle.check_sssp_predecessors(s, d, w, V, got_d.astype(np.float64), got_p, source)
        unreached = np.finfo(wdtype).max
        for v in range(V):
            if got_d[v] == unreached:
                assert got_p[v] == -1
                continue
            cur, steps = v, 0
            while cur != source:
                p = int(got_p[cur])
                assert p >= 0 and got_d[p] <= got_d[cur], (v, cur, p)
                cur, steps = p, steps + 1
                assert steps <= V, f"predecessor cycle reached from {v}"


def test_sssp_cutoff():
    rng = np.random.default_rng(4)
    V, E = 2000, 16000
    s = rng.integers(0, V, E).astype(np.int32)
    d = rng.integers(0, V, E).astype(np.int32)
    w = rng.random(E).astype(np.float32)
    h, g = make_graph(s, d, w, vertices=np.arange(V, dtype=np.int32))
    verts, dist, pred = _sssp(h, g, 0, cutoff=0.7)
    ref_d, _ = oracle.sssp(s, d, w, V, 0, cutoff=0.7)
    assert np.array_equal(by_vertex(verts, dist, V).astype(np.float64), ref_d)


def test_sssp_rmat_symmetric_weighted():
    scale = 16
    s, d = rmat_edgelist(scale, 16 << scale, seed=77)
    rng = np.random.default_rng(2)
    w = rng.random(s.shape[0]).astype(np.float32)
    s2, d2 = _sym(s, d)
    w2 = np.concatenate([w, w])
    V = 1 << scale
    h, g = make_graph(s2, d2, w2, symmetric=True, vertices=np.arange(V, dtype=np.int32))
    src = int(s[0])
    verts, dist, pred = _sssp(h, g, src)
    ref_d, _ = oracle.sssp(s2, d2, w2, V, src)
    got_d = by_vertex(verts, dist, V).astype(np.float64)
    assert np.array_equal(got_d, ref_d)
    assert oracle.check_sssp_predecessors(s2, d2, w2, V, got_d, by_vertex(verts, pred, V), src)


def test_sssp_errors(golden):
    from cugraph_b200 import _capi
    c = golden["c_api"]["sssp_6"]
    h, g = make_graph(c["src"], c["dst"], None)
    with pytest.raises(_capi.CugraphError):
        _sssp(h, g, 0)  # unweighted
    h, g = make_graph(c["src"], c["dst"], c["weights"])
    with pytest.raises(_capi.CugraphError) as e:
        _sssp(h, g, 999)
    assert e.value.code == _capi.INVALID_INPUT


# ------------------------------------------------------------------------------ staging options
def test_symmetrize_and_drop_flags_match_oracle_bfs():
    rng = np.random.default_rng(8)
    V, E = 400, 1500
    s = rng.integers(0, V, E).astype(np.int32)
    d = rng.integers(0, V, E).astype(np.int32)
    h, g = make_graph(s, d, vertices=np.arange(V, dtype=np.int32), symmetrize=True, drop_self_loops=True,
                      drop_multi_edges=True)
    dist, pred, verts = _bfs(h, g, [0], do=True)
    s2, d2 = _sym(s, d)
    ref_d, _ = oracle.bfs(s2, d2, V, [0])
    assert np.array_equal(by_vertex(verts, dist, V), ref_d)


def test_symmetrize_weights_average():
    # (0->1, w=1) and (1->0, w=3) become one undirected edge of weight 2 (symmetrize_edgelist_impl.cuh:92-99)
    src = [0, 1, 1]
    dst = [1, 0, 2]
    w = [1.0, 3.0, 5.0]
    h, g = make_graph(src, dst, w, symmetrize=True)
    verts, dist, _ = _sssp(h, g, 2)
    got = by_vertex(verts, dist, 3)
    assert got.tolist() == [7.0, 5.0, 0.0]


def test_csr_input():
    import torch
    from cugraph_b200 import pylibcugraph as plc
    offs = torch.tensor([0, 1, 3, 6, 7, 8, 8], dtype=torch.int32).cuda()
    idx = torch.tensor([1, 3, 4, 0, 1, 3, 5, 5], dtype=torch.int32).cuda()
    w = torch.tensor([0.1, 2.1, 1.1, 5.1, 3.1, 4.1, 7.2, 3.2], dtype=torch.float32).cuda()
    h = plc.ResourceHandle()
    g = plc.SGGraph(h, plc.GraphProperties(), offs, idx, weight_array=w, input_array_format="CSR", renumber=False)
    verts, dist, pred = plc.sssp(h, g, 0, 10.0, True, False)
    got = by_vertex(verts, dist, 6)
    assert got[1] == pytest.approx(0.1) and got[4] == pytest.approx(1.2) and got[5] == pytest.approx(4.4)
