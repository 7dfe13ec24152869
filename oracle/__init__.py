"""CPU oracle for the PageRank / BFS / SSSP hot path — TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this package.  The product (`cugraph_b200`) never does: it fails loudly when its CUDA
library is missing instead of falling back to anything here.

`oracle.c` restates the reference's own sequential test oracles (file:line cited there); this module
is a thin ctypes/numpy wrapper around it.  Parity is pinned by tests/test_oracle_golden.py against
the reference's golden vectors.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "bench_ref.c")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(_LIB_PATH) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def _p(a, t):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def coo_to_csx(major, minor, num_vertices, weights=None):
    """Counting-sort COO -> (offsets int64[V+1], indices int32[E], weights float64[E]|None)."""
    major = np.ascontiguousarray(major, dtype=np.int32)
    minor = np.ascontiguousarray(minor, dtype=np.int32)
    E = major.shape[0]
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    offsets = np.empty(num_vertices + 1, dtype=np.int64)
    indices = np.empty(E, dtype=np.int32)
    w_out = None if w is None else np.empty(E, dtype=np.float64)
    rc = lib().oracle_coo_to_csx(ctypes.c_int64(E), ctypes.c_int32(num_vertices),
                                 _p(major, ctypes.c_int32), _p(minor, ctypes.c_int32),
                                 _p(w, ctypes.c_double), _p(offsets, ctypes.c_int64),
                                 _p(indices, ctypes.c_int32), _p(w_out, ctypes.c_double))
    if rc != 0:
        raise ValueError(f"oracle_coo_to_csx failed rc={rc}")
    return offsets, indices, w_out


def pagerank(src, dst, num_vertices, weights=None, alpha=0.85, epsilon=1e-5, max_iterations=100,
             precomputed_out_w=None, personalization=None, initial_guess=None, csc=None):
    """fp64 PageRank on a directed edge list (src -> dst). Returns (pr[V], iterations, converged)."""
    if csc is None:
        csc = coo_to_csx(dst, src, num_vertices, weights)
    offsets, indices, w = csc
    pr = np.zeros(num_vertices, dtype=np.float64)
    if initial_guess is not None:
        pr[:] = np.asarray(initial_guess, dtype=np.float64)
    pw = None if precomputed_out_w is None else np.ascontiguousarray(precomputed_out_w, np.float64)
    pv = pvals = None
    psize = 0
    if personalization is not None:
        pv = np.ascontiguousarray(personalization[0], dtype=np.int32)
        pvals = np.ascontiguousarray(personalization[1], dtype=np.float64)
        psize = pv.shape[0]
    it = ctypes.c_int64(0)
    conv = ctypes.c_int(0)
    rc = lib().oracle_pagerank(_p(offsets, ctypes.c_int64), _p(indices, ctypes.c_int32),
                               _p(w, ctypes.c_double), ctypes.c_int32(num_vertices),
                               _p(pw, ctypes.c_double), _p(pv, ctypes.c_int32),
                               _p(pvals, ctypes.c_double), ctypes.c_int32(psize),
                               ctypes.c_int(0 if initial_guess is None else 1),
                               ctypes.c_double(alpha), ctypes.c_double(epsilon),
                               ctypes.c_int64(max_iterations), _p(pr, ctypes.c_double),
                               ctypes.byref(it), ctypes.byref(conv))
    if rc != 0:
        raise ValueError(f"oracle_pagerank failed rc={rc}")
    return pr, int(it.value), bool(conv.value)


def bfs(src, dst, num_vertices, sources, depth_limit=None, csr=None):
    """BFS on a directed edge list. Returns (distances int32[V], predecessors int32[V])."""
    if csr is None:
        csr = coo_to_csx(src, dst, num_vertices)
    offsets, indices, _ = csr
    sources = np.ascontiguousarray(np.atleast_1d(sources), dtype=np.int32)
    dist = np.empty(num_vertices, dtype=np.int32)
    pred = np.empty(num_vertices, dtype=np.int32)
    dl = np.iinfo(np.int32).max if depth_limit is None else int(min(depth_limit, np.iinfo(np.int32).max))
    rc = lib().oracle_bfs(_p(offsets, ctypes.c_int64), _p(indices, ctypes.c_int32),
                          ctypes.c_int32(num_vertices), _p(sources, ctypes.c_int32),
                          ctypes.c_int32(sources.shape[0]), ctypes.c_int32(dl),
                          _p(dist, ctypes.c_int32), _p(pred, ctypes.c_int32))
    if rc != 0:
        raise ValueError(f"oracle_bfs failed rc={rc}")
    return dist, pred


def sssp(src, dst, weights, num_vertices, source, cutoff=None, use_float=True, csr=None):
    """Dijkstra on a directed weighted edge list. Distances returned as float64 holding values
    computed in float32 (use_float) or float64."""
    if csr is None:
        csr = coo_to_csx(src, dst, num_vertices, weights)
    offsets, indices, w = csr
    dist = np.empty(num_vertices, dtype=np.float64)
    pred = np.empty(num_vertices, dtype=np.int32)
    inf = float(np.finfo(np.float32).max) if use_float else float(np.finfo(np.float64).max)
    co = inf if cutoff is None else float(cutoff)
    rc = lib().oracle_sssp(_p(offsets, ctypes.c_int64), _p(indices, ctypes.c_int32),
                           _p(w, ctypes.c_double), ctypes.c_int32(num_vertices),
                           ctypes.c_int32(int(source)), ctypes.c_double(co),
                           ctypes.c_int(1 if use_float else 0), _p(dist, ctypes.c_double),
                           _p(pred, ctypes.c_int32))
    if rc != 0:
        raise ValueError(f"oracle_sssp failed rc={rc}")
    return dist, pred


def spmv_f32(csc, x, alpha, init):
    """One pull SpMV sweep (the CPU-baseline unit)."""
    offsets, indices, w = csc
    V = offsets.shape[0] - 1
    x = np.ascontiguousarray(x, dtype=np.float32)
    w32 = None if w is None else np.ascontiguousarray(w, dtype=np.float32)
    y = np.empty(V, dtype=np.float32)
    lib().oracle_spmv_f32(_p(offsets, ctypes.c_int64), _p(indices, ctypes.c_int32),
                          _p(w32, ctypes.c_float), ctypes.c_int32(V), _p(x, ctypes.c_float),
                          ctypes.c_float(alpha), ctypes.c_float(init), _p(y, ctypes.c_float))
    return y


# ---- the CPU arm of bench.py (bench_ref.c): RMAT on the host, parallel CSC, float32 PageRank on all cores -----------------

def bench_rmat_edges(scale, num_edges, seed=0, a=0.57, b=0.19, c=0.19):
    src = np.empty(num_edges, dtype=np.int32)
    dst = np.empty(num_edges, dtype=np.int32)
    lib().bench_rmat_edges(ctypes.c_int(scale), ctypes.c_int64(num_edges), ctypes.c_uint64(seed), ctypes.c_double(a),
                           ctypes.c_double(b), ctypes.c_double(c), _p(src, ctypes.c_int32), _p(dst, ctypes.c_int32))
    return src, dst


def bench_build_csc(src, dst, num_vertices):
    E = src.shape[0]
    offsets = np.empty(num_vertices + 1, dtype=np.int64)
    indices = np.empty(E, dtype=np.int32)
    out_degree = np.empty(num_vertices, dtype=np.int32)
    rc = lib().bench_build_csc(ctypes.c_int64(E), ctypes.c_int32(num_vertices), _p(src, ctypes.c_int32), _p(dst, ctypes.c_int32),
                               _p(offsets, ctypes.c_int64), _p(indices, ctypes.c_int32), _p(out_degree, ctypes.c_int32))
    if rc != 0:
        raise ValueError(f"bench_build_csc failed rc={rc}")
    return offsets, indices, out_degree


class BenchPageRank:
    """float32 PageRank (bench_ref.c: bench_pagerank_f32) on a prebuilt CSC; run(k) = k more power iterations"""

    def __init__(self, offsets, indices, out_degree, alpha=0.85):
        self.off, self.idx, self.deg, self.alpha = offsets, indices, out_degree, alpha
        self.V = out_degree.shape[0]
        self.pr = np.full(self.V, 1.0 / self.V, dtype=np.float32)
        self.x = np.empty(self.V, dtype=np.float32)
        self.y = np.empty(self.V, dtype=np.float32)

    def reset(self):
        self.pr[:] = 1.0 / self.V

    def run(self, iterations):
        lib().bench_pagerank_f32(_p(self.off, ctypes.c_int64), _p(self.idx, ctypes.c_int32), _p(self.deg, ctypes.c_int32),
                                 ctypes.c_int32(self.V), ctypes.c_double(self.alpha), ctypes.c_int(iterations),
                                 _p(self.pr, ctypes.c_float), _p(self.x, ctypes.c_float), _p(self.y, ctypes.c_float))
        return self.pr


# ---- validity predicates the reference's own tests use -----------------------------------------

def check_bfs_predecessors(src, dst, num_vertices, dist, pred, sources):
    """Tree-validity predicate of cpp/tests/traversal/bfs_test.cpp:213-233: for every reached
    non-source vertex, dist[pred]+1 == dist[v] and edge (pred -> v) exists."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    dist = np.asarray(dist)
    pred = np.asarray(pred)
    inf = np.iinfo(dist.dtype).max
    is_src = np.zeros(num_vertices, dtype=bool)
    is_src[np.atleast_1d(sources)] = True
    reached = dist != inf
    if not np.all(pred[~reached] == -1):
        return False
    if not np.all(pred[is_src] == -1):
        return False
    v = np.nonzero(reached & ~is_src)[0]
    p = pred[v].astype(np.int64)
    if np.any(p < 0) or np.any(dist[p] == inf):
        return False
    if not np.all(dist[p].astype(np.int64) + 1 == dist[v].astype(np.int64)):
        return False
    keys = np.unique(src * num_vertices + dst)
    q = p * num_vertices + v
    pos = np.searchsorted(keys, q)
    pos[pos >= keys.shape[0]] = keys.shape[0] - 1
    return bool(np.all(keys[pos] == q)) if q.size else True


def check_sssp_predecessors(src, dst, weights, num_vertices, dist, pred, source, rel_tol=1e-6):
    """Predicate of cpp/tests/traversal/sssp_test.cpp:222-240: for every reached non-source vertex
    an edge (pred -> v) with dist[pred] + w ~= dist[v] exists."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    w = np.asarray(weights, dtype=np.float64)
    dist = np.asarray(dist, dtype=np.float64)
    pred = np.asarray(pred)
    inf = dist.max() if dist.size else 0
    reached = pred >= 0
    if pred[source] != -1:
        return False
    ok = np.zeros(num_vertices, dtype=bool)
    ok[~reached] = True
    sel = pred[dst] == src
    tol = rel_tol * max(float(w.max()) if w.size else 1.0, 1.0)
    good = sel & (np.abs(dist[src] + w - dist[dst]) <= tol)
    ok[dst[good]] = True
    return bool(np.all(ok))


# ---------------------------------------------------------------------------------------------
# sibling algorithms: numpy restatements of the reference tests' own CPU references
# ---------------------------------------------------------------------------------------------
def katz(src, dst, num_vertices, weights=None, alpha=0.01, beta=1.0, epsilon=1e-6, max_iterations=500, normalize=True,
         dtype=np.float64):
    """katz_centrality_reference (cpp/tests/centrality/katz_centrality_test.cpp:37-103) with the update of
    katz_centrality_impl.cuh:104-196: x <- alpha * A^T x + beta from x = 0, until sum |x_new - x_old| < epsilon;
    then x / ||x||_2.  Arithmetic in `dtype` (the device computes in the graph's weight type).  Returns (x, iterations)."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    w = np.ones(src.size, dtype=np.float64) if weights is None else np.asarray(weights, dtype=np.float64)
    x = np.zeros(num_vertices, dtype=np.float64)
    it = 0
    while True:
        new = np.bincount(dst, weights=alpha * x[src] * w, minlength=num_vertices) + beta
        new = new.astype(dtype).astype(np.float64)
        diff = np.abs(new - x).sum()
        x = new
        it += 1
        if dtype(diff) < dtype(epsilon):
            break
        if it >= max_iterations:
            raise RuntimeError("Katz Centrality failed to converge.")
    if normalize:
        x = x / np.sqrt((x * x).sum())
    return x, it


def hits(src, dst, num_vertices, epsilon=1e-5, max_iterations=500, initial_hubs=None, normalize=True):
    """hits_reference (cpp/tests/link_analysis/hits_test.cpp:40-120) / hits_impl.cuh:49-191: authorities = sum of the
    in-neighbours' hubs, hubs = sum of the out-neighbours' authorities, both divided by their maximum; until
    sum |hubs - previous hubs| < V * epsilon; finally both divided by their sum if `normalize`.  fp64.
    Returns (hubs, authorities, iterations, last difference)."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    V = num_vertices
    if initial_hubs is None:
        prev = np.full(V, 1.0 / V)
    else:
        prev = np.asarray(initial_hubs, dtype=np.float64) / np.sum(initial_hubs)
    it = 0
    while True:
        auth = np.bincount(dst, weights=prev[src], minlength=V)
        curr = np.bincount(src, weights=auth[dst], minlength=V)
        curr = curr / curr.max()
        auth = auth / auth.max()
        diff = np.abs(curr - prev).sum()
        prev = curr
        it += 1
        if diff < V * epsilon:
            break
        if it >= max_iterations:
            raise RuntimeError("HITS failed to converge.")
    if normalize:
        prev = prev / prev.sum()
        auth = auth / auth.sum()
    return prev, auth, it, diff


def wcc(src, dst, num_vertices):
    """component index per vertex of the undirected graph — the role of weakly_connected_components_reference
    (cpp/tests/components/weakly_connected_components_test.cpp:36-77: BFS from every unvisited vertex): plain union-find."""
    parent = np.arange(num_vertices, dtype=np.int64)

    def find(v):
        while parent[v] != v:
            parent[v] = parent[parent[v]]
            v = parent[v]
        return v

    for u, v in zip(np.asarray(src).tolist(), np.asarray(dst).tolist()):
        ru, rv = find(u), find(v)
        if ru != rv:
            parent[max(ru, rv)] = min(ru, rv)
    return np.array([find(v) for v in range(num_vertices)], dtype=np.int64)


def eigenvector(src, dst, num_vertices, weights=None, epsilon=1e-6, max_iterations=500):
    """eigenvector_centrality_reference (cpp/tests/centrality/eigenvector_centrality_test.cpp:37-100) /
    eigenvector_centrality_impl.cuh:34-150: x <- (A^T x + x) / ||A^T x + x||_2 from x = 1 / V, until
    sum |x_new - x_old| < V * epsilon.  fp64.  Returns (x, iterations)."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    w = np.ones(src.size) if weights is None else np.asarray(weights, dtype=np.float64)
    x = np.full(num_vertices, 1.0 / num_vertices)
    it = 0
    while True:
        new = np.bincount(dst, weights=x[src] * w, minlength=num_vertices) + x
        new = new / np.sqrt((new * new).sum())
        diff = np.abs(new - x).sum()
        x = new
        it += 1
        if diff < num_vertices * epsilon:
            break
        if it >= max_iterations:
            raise RuntimeError("Eigenvector Centrality failed to converge.")
    return x, it
