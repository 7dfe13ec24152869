"""BASELINE.json's full-size configurations (RMAT scale-24 ef-16) checked through size-independent certificates, computed
with plain torch on the device — the CPU oracle would take minutes at this size:

* PageRank: total mass 1, and ONE iteration started from the library's own 5-iteration vector (initial_guess) equals an
  independent fp64 torch evaluation of the reference's update rule (pagerank_impl.cuh:256-283) within 1e-6 relative —
  every edge of the 2^28-edge sweep contributes to that comparison;
* BFS (direction-optimising): distances are THE shortest hop counts iff dist[source] = 0, no edge (u, v) has
  dist[v] > dist[u] + 1, and every other reached vertex has a predecessor edge with dist[pred] + 1 == dist[v]
  (bfs_test.cpp:213-233 is this predicate) — checked over all 2^29 directed edges: bit-exact without an oracle;
* SSSP: dist[source] = 0, no edge can still be relaxed (dist[v] <= fl(dist[u] + w)) and every other reached vertex has a
  predecessor edge that attains its distance exactly (sssp_test.cpp:222-240) — the fixpoint characterisation of the same
  float arithmetic.

CUGRAPH_B200_FULL_SCALE overrides the scale (the CPU emulation run of this file uses a small one)."""
import os

import pytest

pytestmark = pytest.mark.gpu

SCALE = int(os.environ.get("CUGRAPH_B200_FULL_SCALE", "24"))
INT_MAX = 2**31 - 1
FLT_MAX = 3.4028234663852886e38


def _edges():
    from cugraph_b200.generators import rmat_edgelist
    return rmat_edgelist(SCALE, 16 << SCALE, seed=0)


def _positions(verts, n_ids):
    """external id -> index into the result arrays (-1: not a vertex of the graph)"""
    import torch
    pos = torch.full((n_ids,), -1, dtype=torch.int64, device=verts.device)
    pos[verts.long()] = torch.arange(verts.numel(), device=verts.device)
    return pos


def test_pagerank_full_size_one_step_certificate():
    import torch
    from cugraph_b200 import pylibcugraph as plc
    alpha = 0.85
    src, dst = _edges()
    h = plc.ResourceHandle()
    g = plc.SGGraph(h, plc.GraphProperties(is_multigraph=True), src, dst, store_transposed=True, renumber=True)
    v5, p5, _ = plc.pagerank(h, g, None, None, None, None, alpha, 0.0, 5, False, fail_on_nonconvergence=False)
    v6, p6, _ = plc.pagerank(h, g, None, None, v5, p5, alpha, 0.0, 1, False, fail_on_nonconvergence=False)
    assert torch.equal(v5, v6)
    n = v5.numel()
    assert abs(float(p5.double().sum()) - 1.0) < 1e-5 and abs(float(p6.double().sum()) - 1.0) < 1e-5
    pos = _positions(v5, 1 << SCALE)
    sp, dp = pos[src.long()], pos[dst.long()]
    assert int(sp.min()) >= 0 and int(dp.min()) >= 0          # every endpoint is a vertex of the graph
    outdeg = torch.bincount(sp, minlength=n).double()
    p0 = p5.double()
    p0 = p0 / p0.sum()                                          # the driver normalises an initial guess (pagerank_impl.cuh:196-211)
    x = torch.where(outdeg > 0, p0 / outdeg.clamp(min=1.0), torch.zeros_like(p0))
    dangling = p0[outdeg == 0].sum()
    y = torch.zeros(n, dtype=torch.float64, device=p0.device).index_add_(0, dp, x[sp])
    ref = y * alpha + (alpha * dangling + (1.0 - alpha)) / n
    rel = ((p6.double() - ref).abs() / ref).max()
    assert float(rel) < 1e-6, float(rel)
    # 100 iterations (the bench step): mass stays 1
    _, p100, conv = plc.pagerank(h, g, None, None, None, None, alpha, 0.0, 100, False, fail_on_nonconvergence=False)
    assert abs(float(p100.double().sum()) - 1.0) < 1e-5 and not conv


def _sym_graph(weighted):
    import torch
    from cugraph_b200 import pylibcugraph as plc
    src, dst = _edges()
    s2, d2 = torch.cat([src, dst]), torch.cat([dst, src])
    del src, dst
    w2 = None
    if weighted:
        gen = torch.Generator(device="cuda")
        gen.manual_seed(2)
        w = torch.rand(s2.numel() // 2, device="cuda", generator=gen)
        w2 = torch.cat([w, w])
    h = plc.ResourceHandle()
    g = plc.SGGraph(h, plc.GraphProperties(is_symmetric=True, is_multigraph=True), s2, d2, weight_array=w2,
                    store_transposed=False, renumber=True)
    return h, g, s2, d2, w2


def _a_source(s2):
    import torch
    deg = torch.bincount(s2.long(), minlength=1 << SCALE)
    cand = torch.nonzero(deg > 0).flatten()
    torch.manual_seed(1)
    return int(cand[torch.randint(0, cand.numel(), (1,), device=cand.device)].item())


def test_bfs_full_size_certificate():
    import torch
    from cugraph_b200 import pylibcugraph as plc
    h, g, s2, d2, _ = _sym_graph(False)
    source = _a_source(s2)
    dist, pred, verts = plc.bfs(h, g, torch.tensor([source], dtype=torch.int32, device="cuda"), True, 0, True, False)
    pos = _positions(verts, 1 << SCALE)
    sp, dp = pos[s2.long()], pos[d2.long()]
    du, dv = dist[sp].long(), dist[dp].long()
    assert int(dist[pos[source]]) == 0
    assert bool((dv <= du + 1).all())                          # no edge skips a level (INT_MAX + 1 does not wrap in int64)
    reached = dist != INT_MAX
    has_pred = pred >= 0
    assert bool((has_pred == (reached & (verts != source))).all())
    assert bool((dist[pos[pred[has_pred].long()]] + 1 == dist[has_pred]).all())
    # the predecessor is a neighbour: some edge (pred[v], v) exists
    hit = torch.zeros(verts.numel(), dtype=torch.bool, device=verts.device)
    hit[dp[pred[dp] == s2]] = True
    assert bool((hit == has_pred).all())
    assert int(reached.sum()) > verts.numel() // 2             # the giant component


def test_sssp_full_size_certificate():
    import torch
    from cugraph_b200 import pylibcugraph as plc
    h, g, s2, d2, w2 = _sym_graph(True)
    source = _a_source(s2)
    verts, dist, pred = plc.sssp(h, g, source, float("inf"), True, False)
    assert dist.dtype == torch.float32
    pos = _positions(verts, 1 << SCALE)
    sp, dp = pos[s2.long()], pos[d2.long()]
    assert float(dist[pos[source]]) == 0.0
    reached = dist < FLT_MAX
    du, dv = dist[sp], dist[dp]
    cand = du + w2                                              # float32 add, the library's own arithmetic
    ok = reached[sp]
    assert bool((dv[ok] <= cand[ok]).all())                     # nothing left to relax
    assert bool((reached[dp][ok]).all())                        # neighbours of reached vertices are reached
    has_pred = pred >= 0
    assert bool((has_pred == (reached & (verts != source))).all())
    # the predecessor edge attains the distance exactly
    tree = ok & (pred[dp] == s2)
    best = torch.full((verts.numel(),), float("inf"), dtype=torch.float32, device=verts.device)
    best.scatter_reduce_(0, dp[tree], cand[tree], reduce="amin")
    assert bool((best[has_pred] == dist[has_pred]).all())
