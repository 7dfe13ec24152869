"""Pin the CPU oracle (oracle/oracle.c) against every golden vector the reference's own tests hold
for this path.  CPU only."""
import numpy as np
import pytest

import oracle

FLT_MAX = float(np.finfo(np.float32).max)


def _nearly_equal(a, b, eps):
    # nearlyEqual of cpp/tests/c_api/test_utils.cpp: relative to the larger magnitude
    return abs(a - b) <= max(abs(a), abs(b)) * eps


@pytest.mark.parametrize("case", ["pagerank_6", "pagerank_6_nonconverged", "pagerank_4",
                                  "personalized_pagerank_4", "personalized_pagerank_4_nonconverged"])
def test_pagerank_c_api_golden(golden, case):
    g = golden["c_api"][case]
    pers = None
    if "personalization_vertices" in g:
        pers = (g["personalization_vertices"], g["personalization_values"])
    w = np.asarray(g["weights"], dtype=np.float32)
    pr, iters, conv = oracle.pagerank(g["src"], g["dst"], g["num_vertices"], w, alpha=g["alpha"],
                                      epsilon=g["epsilon"], max_iterations=g["max_iterations"],
                                      personalization=pers)
    for a, b in zip(pr, g["values"]):
        assert _nearly_equal(a, b, g["rel_tol"])
    assert conv == ("nonconverged" not in case)


@pytest.mark.parametrize("name", ["karate.csv", "dolphins.csv", "Simple_1", "Simple_2"])
def test_pagerank_pylibcugraph_golden(golden, name):
    g = golden["pylibcugraph"][name]
    p = g["pagerank"]
    nv = len(p["vertices"])
    pr, iters, conv = oracle.pagerank(g["src"], g["dst"], nv, np.asarray(g["weights"], np.float32),
                                      alpha=p["alpha"], epsilon=p["epsilon"],
                                      max_iterations=p["max_iterations"])
    assert conv
    np.testing.assert_allclose(pr, p["values"], rtol=p["rel_tol"])


def test_bfs_c_api_golden(golden):
    g = golden["c_api"]["bfs_6"]
    dist, pred = oracle.bfs(g["src"], g["dst"], g["num_vertices"], g["sources"], g["depth_limit"])
    assert dist.tolist() == g["distances"]
    assert pred.tolist() == g["predecessors"]
    assert oracle.check_bfs_predecessors(g["src"], g["dst"], g["num_vertices"], dist, pred, g["sources"])


def test_bfs_depth_limit():
    src = [0, 1, 2, 3]
    dst = [1, 2, 3, 4]
    dist, pred = oracle.bfs(src, dst, 5, [0], depth_limit=2)
    assert dist.tolist() == [0, 1, 2, 2147483647, 2147483647]
    assert pred.tolist() == [-1, 0, 1, -1, -1]


@pytest.mark.parametrize("use_float", [True, False])
def test_sssp_c_api_golden(golden, use_float):
    g = golden["c_api"]["sssp_6"]
    w = np.asarray(g["weights"], dtype=np.float32 if use_float else np.float64)
    dist, pred = oracle.sssp(g["src"], g["dst"], w, g["num_vertices"], g["source"], g["cutoff"], use_float)
    big = FLT_MAX if use_float else float(np.finfo(np.float64).max)
    exp = [big if d == "MAX" else d for d in g["distances"]]
    for a, b in zip(dist, exp):
        assert _nearly_equal(a, b, 1e-3)
    assert pred.tolist() == g["predecessors"]


@pytest.mark.parametrize("name", ["karate.csv", "dolphins.csv", "Simple_1", "Simple_2"])
def test_sssp_pylibcugraph_golden(golden, name):
    g = golden["pylibcugraph"][name]
    s = g["sssp"]
    nv = len(s["distances"])
    w = np.asarray(g["weights"], dtype=np.float32)
    dist, pred = oracle.sssp(g["src"], g["dst"], w, nv, s["source"], s["cutoff"], True)
    for a, b in zip(dist, s["distances"]):
        if a <= 3.4e38 or b <= 3.4e38:
            assert a == pytest.approx(b, 1e-4)
    if s["predecessors_checked"]:
        assert pred.tolist() == s["predecessors"]
    assert oracle.check_sssp_predecessors(g["src"], g["dst"], w, nv, dist, pred, s["source"])


def test_karate_vs_networkx(golden):
    """The reference's NetworkX protocol (python/cugraph/.../test_pagerank.py:77-105,190-200):
    NetworkX at tol*0.01, 2x iterations; < 1 % of vertices may differ by more than 1.1*tol."""
    nx = pytest.importorskip("networkx")
    g = golden["pylibcugraph"]["karate.csv"]
    tol = 1e-5
    G = nx.DiGraph()
    G.add_weighted_edges_from(zip(g["src"], g["dst"], g["weights"]))
    ref = nx.pagerank(G, alpha=0.85, tol=tol * 0.01, max_iter=200)
    pr, _, conv = oracle.pagerank(g["src"], g["dst"], 34, np.asarray(g["weights"], np.float32),
                                  alpha=0.85, epsilon=tol, max_iterations=100)
    assert conv
    bad = sum(abs(pr[v] - ref[v]) > tol * 1.1 for v in range(34))
    assert bad < 0.01 * 34


def test_coo_to_csx_sorted_rows():
    rng = np.random.default_rng(0)
    V, E = 50, 600
    s = rng.integers(0, V, E).astype(np.int32)
    d = rng.integers(0, V, E).astype(np.int32)
    w = rng.random(E)
    off, idx, wo = oracle.coo_to_csx(s, d, V, w)
    assert off[0] == 0 and off[-1] == E
    for v in range(V):
        row = idx[off[v]:off[v + 1]]
        assert np.all(np.diff(row) >= 0)
        assert sorted(row.tolist()) == sorted(d[s == v].tolist())
    assert np.isclose(wo.sum(), w.sum())
