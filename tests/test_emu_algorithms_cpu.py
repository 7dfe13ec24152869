"""PageRank, BFS and SSSP on the CPU: the single-GPU sources (pagerank.cu, traverse.cu, spmv*.cuh, graph_build.cu) compiled
as plain C++ against the SIMT emulation in emu/ (threads of a CTA are fibers; warp collectives and __syncthreads are
real rendezvous points; TMA copies are memcpys) and called through the C ABI with numpy arrays — kernel LOGIC (work
distribution, stealing, segmented reductions, frontier queues, device-resident loop state) checked against the oracle
without a GPU.  Performance, memory-model and scheduling effects are of course only visible on the GPU (`-m gpu`)."""
import ctypes as C

import numpy as np
import pytest

import oracle
from tests.test_emu_staging_cpu import FLOAT32, INT32, Props, create_graph, emu, make_edges  # noqa: F401


def _view_to_np(L, view):
    L.cugraph_type_erased_device_array_view_size.restype = C.c_size_t
    L.cugraph_type_erased_device_array_view_size.argtypes = [C.c_void_p]
    L.cugraph_type_erased_device_array_view_type.restype = C.c_int
    L.cugraph_type_erased_device_array_view_type.argtypes = [C.c_void_p]
    L.cugraph_type_erased_device_array_view_pointer.restype = C.c_void_p
    L.cugraph_type_erased_device_array_view_pointer.argtypes = [C.c_void_p]
    n = L.cugraph_type_erased_device_array_view_size(view)
    t = L.cugraph_type_erased_device_array_view_type(view)
    dt = {2: np.int32, 3: np.int64, 8: np.float32, 9: np.float64}[t]
    p = L.cugraph_type_erased_device_array_view_pointer(view)
    out = np.ctypeslib.as_array(C.cast(p, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy() if n else np.zeros(0, dt)
    L.cugraph_type_erased_device_array_view_free(C.c_void_p(view))
    return out


def run_pagerank(L, g, alpha, eps, max_it):
    res, err = C.c_void_p(), C.c_void_p()
    code = L.cugraph_pagerank_allow_nonconvergence(C.c_void_p(L.handle), g, None, None, None, None, C.c_double(alpha),
                                                   C.c_double(eps), C.c_size_t(max_it), 0, C.byref(res), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    for f in ("cugraph_centrality_result_get_vertices", "cugraph_centrality_result_get_values"):
        getattr(L, f).restype = C.c_void_p
        getattr(L, f).argtypes = [C.c_void_p]
    L.cugraph_centrality_result_get_num_iterations.restype = C.c_size_t
    L.cugraph_centrality_result_get_num_iterations.argtypes = [C.c_void_p]
    L.cugraph_centrality_result_free.argtypes = [C.c_void_p]
    v = _view_to_np(L, L.cugraph_centrality_result_get_vertices(res))
    p = _view_to_np(L, L.cugraph_centrality_result_get_values(res))
    it = L.cugraph_centrality_result_get_num_iterations(res)
    L.cugraph_centrality_result_free(res)
    return v, p, int(it)


def dense_ids(src, dst):
    ids, inv = np.unique(np.concatenate([src, dst]), return_inverse=True)
    return ids, inv[:src.size].astype(np.int32), inv[src.size:].astype(np.int32)


PR_CASES = [("plain sweep", "1000000000", False), ("piece-stream sweep", "0", False), ("piece-stream sweep weighted", "0", True)]


@pytest.mark.parametrize("name,min_edges,weighted", PR_CASES, ids=[c[0] for c in PR_CASES])
def test_pagerank_emulated(emu, monkeypatch, name, min_edges, weighted):  # noqa: F811
    monkeypatch.setenv("CUGRAPH_B200_SWEEP_MIN_EDGES", min_edges)
    src, dst, w = make_edges(60_000, 250_000, seed=41, weighted=weighted, id_offset=5)
    g = create_graph(emu, src, dst, w)
    verts, pr, it = run_pagerank(emu, g, 0.85, 0.0, 20)             # equal iteration counts: the parity protocol
    ids, s, d = dense_ids(src, dst)
    ref, ref_it, _ = oracle.pagerank(s, d, ids.size, None if w is None else w.astype(np.float64), alpha=0.85, epsilon=0.0,
                                     max_iterations=20)
    assert it == ref_it == 20
    got = np.zeros(ids.size)
    got[np.searchsorted(ids, verts)] = pr
    np.testing.assert_allclose(got, ref, rtol=2e-5 if weighted else 1e-5, atol=0)
    assert abs(got.sum() - 1.0) < 1e-5
    emu.cugraph_graph_free(g)


def test_pagerank_emulated_stealing_and_tiny_phases(emu, monkeypatch):  # noqa: F811
    """sources spread over many column blocks with few chunks each: under emulation CTA 0 runs its own phases and then
    drains every other CTA's phases through the stealing path (CTAs run one after the other)"""
    monkeypatch.setenv("CUGRAPH_B200_SWEEP_MIN_EDGES", "0")
    src, dst, w = make_edges(400_000, 500_000, seed=43)
    g = create_graph(emu, src, dst, w)
    verts, pr, it = run_pagerank(emu, g, 0.85, 0.0, 6)
    ids, s, d = dense_ids(src, dst)
    ref, _, _ = oracle.pagerank(s, d, ids.size, None, alpha=0.85, epsilon=0.0, max_iterations=6)
    got = np.zeros(ids.size)
    got[np.searchsorted(ids, verts)] = pr
    assert it == 6
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=0)
    emu.cugraph_graph_free(g)


@pytest.mark.parametrize("weighted", [False, True])
def test_sweep_against_plain_sweep_emulated(emu, monkeypatch, weighted):  # noqa: F811
    """cugraph_b200_debug_compare_sweeps: the piece-stream sweep against the plain sweep (independent kernels) row by row"""
    monkeypatch.setenv("CUGRAPH_B200_SWEEP_MIN_EDGES", "0")
    src, dst, w = make_edges(150_000, 700_000, seed=61 + weighted, weighted=weighted)
    g = create_graph(emu, src, dst, w)
    out = (C.c_double * 8)()
    err = C.c_void_p()
    code = emu.cugraph_b200_debug_compare_sweeps(C.c_void_p(emu.handle), g, out, C.byref(err))
    assert code == 0, emu.cugraph_error_message(err)
    assert out[0] < 2e-6 and out[4] < 2e-6 and out[3] == 0 and out[7] == 0, list(out)
    emu.cugraph_graph_free(g)


def test_pagerank_emulated_double_weights(emu, monkeypatch):  # noqa: F811
    """fp64 graphs: 24,512 columns per shared-memory slice, several blocks"""
    from tests.test_emu_staging_cpu import FLOAT64
    monkeypatch.setenv("CUGRAPH_B200_SWEEP_MIN_EDGES", "0")
    src, dst, w32 = make_edges(80_000, 300_000, seed=47, weighted=True)
    w = w32.astype(np.float64) * 1.000000123
    L = emu
    views = [L.cugraph_type_erased_device_array_view_create(a.ctypes.data, a.size, t) for a, t in ((src, INT32), (dst, INT32), (w, FLOAT64))]
    g, err = C.c_void_p(), C.c_void_p()
    code = L.cugraph_graph_create_with_times_sg(C.c_void_p(L.handle), C.byref(Props(0, 1)), None, C.c_void_p(views[0]),
                                                C.c_void_p(views[1]), C.c_void_p(views[2]), None, None, None, None, 1, 1, 0, 0, 0, 0,
                                                C.byref(g), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    verts, pr, it = run_pagerank(L, g, 0.85, 0.0, 12)
    assert pr.dtype == np.float64
    ids, s, d = dense_ids(src, dst)
    ref, _, _ = oracle.pagerank(s, d, ids.size, w, alpha=0.85, epsilon=0.0, max_iterations=12)
    got = np.zeros(ids.size)
    got[np.searchsorted(ids, verts)] = pr
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=0)
    L.cugraph_graph_free(g)


def _paths(L, res):
    for f in ("cugraph_paths_result_get_vertices", "cugraph_paths_result_get_distances", "cugraph_paths_result_get_predecessors"):
        getattr(L, f).restype = C.c_void_p
        getattr(L, f).argtypes = [C.c_void_p]
    L.cugraph_paths_result_free.argtypes = [C.c_void_p]
    v = _view_to_np(L, L.cugraph_paths_result_get_vertices(res))
    dist = _view_to_np(L, L.cugraph_paths_result_get_distances(res))
    pred = _view_to_np(L, L.cugraph_paths_result_get_predecessors(res))
    L.cugraph_paths_result_free(res)
    return v, dist, pred


def symmetric_edges(V, E, seed):
    src, dst, _ = make_edges(V, E, seed)
    s = np.concatenate([src, dst]).astype(np.int32)
    d = np.concatenate([dst, src]).astype(np.int32)
    return s, d


def create_sym_graph(L, s, d, w):
    views = [L.cugraph_type_erased_device_array_view_create(a.ctypes.data, a.size, t) if a is not None else None
             for a, t in ((s, INT32), (d, INT32), (w, 9 if (w is not None and w.dtype == np.float64) else FLOAT32))]  # 9 = FLOAT64
    g, err = C.c_void_p(), C.c_void_p()
    code = L.cugraph_graph_create_with_times_sg(
        C.c_void_p(L.handle), C.byref(Props(1, 1)), None, C.c_void_p(views[0]), C.c_void_p(views[1]),
        C.c_void_p(views[2]) if views[2] else None, None, None, None, None, 0, 1, 0, 0, 0, 0, C.byref(g), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    for v in views:
        if v:
            L.cugraph_type_erased_device_array_view_free(v)
    return g


@pytest.mark.parametrize("direction_optimizing", [0, 1])
def test_bfs_emulated(emu, direction_optimizing):  # noqa: F811
    s, d = symmetric_edges(20_000, 120_000, seed=51)
    g = create_sym_graph(emu, s, d, None)
    ids, ss, dd = dense_ids(s, d)
    source = int(ids[np.bincount(ss).argmax()])                   # a hub: the bottom-up switch triggers
    srcs = np.array([source], dtype=np.int32)
    sv = emu.cugraph_type_erased_device_array_view_create(srcs.ctypes.data, 1, INT32)
    res, err = C.c_void_p(), C.c_void_p()
    code = emu.cugraph_bfs(C.c_void_p(emu.handle), g, C.c_void_p(sv), direction_optimizing, C.c_size_t(2**31 - 2), 1, 0,
                           C.byref(res), C.byref(err))
    assert code == 0, emu.cugraph_error_message(err)
    verts, dist, pred = _paths(emu, res)
    ref_d, _ = oracle.bfs(ss, dd, ids.size, [int(np.searchsorted(ids, source))])
    got = np.zeros(ids.size, dtype=np.int64)
    got[np.searchsorted(ids, verts)] = dist
    unreached = ref_d < 0 if (ref_d < 0).any() else ref_d == np.iinfo(ref_d.dtype).max
    assert (got[~unreached] == ref_d[~unreached]).all()
    assert (got[unreached] == np.iinfo(np.int32).max).all()
    gp = np.zeros(ids.size, dtype=np.int64)
    gp[np.searchsorted(ids, verts)] = pred
    has = (~unreached) & (got > 0)
    pidx = np.searchsorted(ids, gp[has])
    assert (got[pidx] + 1 == got[has]).all()                      # every predecessor is one level closer
    emu.cugraph_graph_free(g)


def test_sssp_emulated(emu):  # noqa: F811
    s, d = symmetric_edges(15_000, 90_000, seed=61)
    r = np.random.default_rng(8)
    half = s.size // 2
    wh = (r.random(half).astype(np.float32) + 0.01)
    w = np.concatenate([wh, wh])                                   # symmetric weights
    g = create_sym_graph(emu, s, d, w)
    ids, ss, dd = dense_ids(s, d)
    source = int(ids[7])
    res, err = C.c_void_p(), C.c_void_p()
    emu.cugraph_sssp.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    code = emu.cugraph_sssp(C.c_void_p(emu.handle), g, source, float("inf"), 1, 0, C.byref(res), C.byref(err))
    assert code == 0, emu.cugraph_error_message(err)
    verts, dist, pred = _paths(emu, res)
    ref_d, _ = oracle.sssp(ss, dd, w, ids.size, int(np.searchsorted(ids, source)), use_float=True)
    got = np.zeros(ids.size, dtype=np.float32)
    got[np.searchsorted(ids, verts)] = dist
    assert (got == ref_d.astype(np.float32)).all()                 # bit-exact in float, unreached = FLT_MAX on both sides
    emu.cugraph_graph_free(g)


def _sssp_dist(emu, g, source):  # noqa: F811
    res, err = C.c_void_p(), C.c_void_p()
    emu.cugraph_sssp.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    code = emu.cugraph_sssp(C.c_void_p(emu.handle), g, source, float("inf"), 1, 0, C.byref(res), C.byref(err))
    assert code == 0, emu.cugraph_error_message(err)
    return _paths(emu, res)


def follow_to_source(verts, dist, pred, source, unreached):
    """every reached vertex's predecessor chain must end at the source (no cycles), distances never increasing along it"""
    pos = {int(v): i for i, v in enumerate(verts)}
    for i, v in enumerate(verts):
        if dist[i] == unreached:
            assert pred[i] == -1
            continue
        cur, steps = int(v), 0
        while cur != source:
            j = pos[cur]
            p = int(pred[j])
            assert p >= 0, f"vertex {cur} (reached from {int(v)}) has no predecessor"
            assert dist[pos[p]] <= dist[j]
            cur = p
            steps += 1
            assert steps <= len(verts), f"predecessor cycle reached from vertex {int(v)}"


@pytest.mark.parametrize("wdtype", [np.float32, np.float64])
def test_sssp_zero_weight_predecessors_emulated(emu, wdtype):  # noqa: F811
    """symmetric zero-weight edges and zero-weight cycles: the distance fixpoint alone cannot orient them (both directions of
    the edge are tight), the predecessors must still form a tree rooted at the source (float32: packed word at the relaxation,
    float64: strict pass + tie passes)"""
    r = np.random.default_rng(3)
    V = 400
    half_s = r.integers(0, V, 1600).astype(np.int32)
    half_d = r.integers(0, V, 1600).astype(np.int32)
    wh = np.where(r.random(1600) < 0.5, 0.0, r.random(1600)).astype(np.float32)      # half of the edges weigh nothing
    # a zero-weight cycle 7 - 8 - 9 - 7 hanging off vertex 6, and a float-absorbed edge (1e8 + 1 == 1e8)
    extra = [(6, 7, 0.0), (7, 8, 0.0), (8, 9, 0.0), (9, 7, 0.0), (0, 390, 1e8), (390, 391, 1.0), (391, 392, 1.0)]
    half_s = np.concatenate([half_s, np.array([e[0] for e in extra], np.int32)])
    half_d = np.concatenate([half_d, np.array([e[1] for e in extra], np.int32)])
    wh = np.concatenate([wh, np.array([e[2] for e in extra], np.float32)])
    if wdtype == np.float64:
        wh = wh.astype(np.float64)
        wh[wh == 1e8] = 1e16
    s, d, w = np.concatenate([half_s, half_d]), np.concatenate([half_d, half_s]), np.concatenate([wh, wh])
    g = create_sym_graph(emu, s, d, w)
    ids, ss, dd = dense_ids(s, d)
    for source in (int(ids[0]), int(ids[7])):
        verts, dist, pred = _sssp_dist(emu, g, source)
        ref_d, _ = oracle.sssp(ss, dd, w, ids.size, int(np.searchsorted(ids, source)), use_float=(wdtype == np.float32))
        got = np.zeros(ids.size, dtype=wdtype)
        got[np.searchsorted(ids, verts)] = dist
        assert (got == ref_d.astype(wdtype)).all()
        follow_to_source(verts, dist, pred, source, np.finfo(wdtype).max)
        assert oracle.check_sssp_predecessors(ss, dd, w, ids.size, ref_d, _scatter(ids, verts, pred), int(np.searchsorted(ids, source)))
    emu.cugraph_graph_free(g)


def _scatter(ids, verts, pred):
    """predecessors (external ids, -1 = none) as an array over dense vertex numbers holding dense predecessor numbers"""
    out = np.full(ids.size, -1, dtype=np.int32)
    has = pred >= 0
    out[np.searchsorted(ids, verts[has])] = np.searchsorted(ids, pred[has])
    return out


@pytest.mark.parametrize("weights", ["uniform", "tiny-and-huge", "constant"])
def test_sssp_window_control_emulated(emu, monkeypatch, capfd, weights):  # noqa: F811
    """The window-width controller (64x narrower start, doubling / halving by rounds, mid-window split of a busy window)
    only schedules the work: distances and predecessors' validity do not depend on it.  Hub-heavy graph so that splits
    happen; weight sets that stress the float arithmetic of the window bounds."""
    s, d = symmetric_edges(12_000, 150_000, seed=5)
    r = np.random.default_rng(9)
    half = s.size // 2
    if weights == "uniform":
        wh = r.random(half).astype(np.float32)
    elif weights == "tiny-and-huge":
        wh = np.where(r.random(half) < 0.5, 1e-6, 1e6).astype(np.float32) * (1.0 + r.random(half).astype(np.float32))
    else:
        wh = np.full(half, 0.25, dtype=np.float32)
    w = np.concatenate([wh, wh])
    g = create_sym_graph(emu, s, d, w)
    ids, ss, dd = dense_ids(s, d)
    source = int(ids[3])
    ref_d, _ = oracle.sssp(ss, dd, w, ids.size, int(np.searchsorted(ids, source)), use_float=True)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("CUGRAPH_B200_SSSP_ADAPTIVE", mode)
        monkeypatch.setenv("CUGRAPH_B200_SSSP_TRACE", "1")
        monkeypatch.setenv("CUGRAPH_B200_SSSP_SPLIT_MIN_EDGES", "0")   # the default only splits rounds of >= 2^20 edges
        emu.emu_reload_tuning(C.c_void_p(emu.handle))                   # knobs are read per handle
        capfd.readouterr()
        verts, dist, pred = _sssp_dist(emu, g, source)
        trace = capfd.readouterr().err
        got = np.zeros(ids.size, dtype=np.float32)
        got[np.searchsorted(ids, verts)] = dist
        assert (got == ref_d.astype(np.float32)).all(), mode
        out[mode] = trace
    last = [ln for ln in out["1"].splitlines() if ln.startswith("sssp window")][-1]
    if weights == "uniform":
        assert int(last.split("splits so far")[1]) > 0, last       # the controller did cut a busy window
    emu.cugraph_graph_free(g)


def test_advance_in_halves_emulated(emu, monkeypatch):  # noqa: F811
    """a frontier whose degree sum reaches the 32-bit tile numbering's limit is advanced in halves (graphs with 64-bit
    offsets); the limit is lowered to 500 edges here so that BFS and SSSP take that path at every level"""
    monkeypatch.setenv("CUGRAPH_B200_ADVANCE_SPLIT_EDGES", "500")
    emu.emu_reload_tuning(C.c_void_p(emu.handle))
    try:
        s, d = symmetric_edges(4_000, 30_000, seed=13)
        w = np.random.default_rng(4).random(s.size // 2).astype(np.float32)
        w = np.concatenate([w, w])
        g = create_sym_graph(emu, s, d, w)
        ids, ss, dd = dense_ids(s, d)
        source = int(ids[5])
        srcs = np.array([source], dtype=np.int32)
        sv = emu.cugraph_type_erased_device_array_view_create(srcs.ctypes.data, 1, INT32)
        res, err = C.c_void_p(), C.c_void_p()
        code = emu.cugraph_bfs(C.c_void_p(emu.handle), g, C.c_void_p(sv), 0, C.c_size_t(2**31 - 2), 1, 0, C.byref(res), C.byref(err))
        assert code == 0, emu.cugraph_error_message(err)
        verts, dist, _ = _paths(emu, res)
        ref_d, _ = oracle.bfs(ss, dd, ids.size, [int(np.searchsorted(ids, source))])
        got = np.zeros(ids.size, dtype=np.int64)
        got[np.searchsorted(ids, verts)] = dist
        ref = np.asarray(ref_d, dtype=np.int64)
        reached = (ref >= 0) & (ref < np.iinfo(np.int32).max)
        assert (got[reached] == ref[reached]).all() and (got[~reached] == np.iinfo(np.int32).max).all()
        verts, sd, _ = _sssp_dist(emu, g, source)
        ref_s, _ = oracle.sssp(ss, dd, w, ids.size, int(np.searchsorted(ids, source)), use_float=True)
        gs = np.zeros(ids.size, dtype=np.float32)
        gs[np.searchsorted(ids, verts)] = sd
        assert (gs == ref_s.astype(np.float32)).all()
        emu.cugraph_graph_free(g)
    finally:
        monkeypatch.delenv("CUGRAPH_B200_ADVANCE_SPLIT_EDGES")
        emu.emu_reload_tuning(C.c_void_p(emu.handle))


def test_extract_paths_emulated(emu):  # noqa: F811
    """cugraph_extract_paths on a BFS result: every row is the path source ... destination (consecutive vertices are
    edges, length = distance + 1), -1 behind it; unreachable destinations give an all -1 row; the source gives a row of one.
    (The reference's own extract_paths_test.c runs against the library in tests/test_reference_c_tests_*.py.)"""
    L = emu
    s, d = symmetric_edges(5_000, 9_000, seed=77)                  # sparse: several components, deep BFS
    g = create_sym_graph(L, s, d, None)
    ids, ss, dd = dense_ids(s, d)
    source = int(ids[np.bincount(ss).argmax()])
    srcs = np.array([source], dtype=np.int32)
    sv = L.cugraph_type_erased_device_array_view_create(srcs.ctypes.data, 1, INT32)
    res, err = C.c_void_p(), C.c_void_p()
    code = L.cugraph_bfs(C.c_void_p(L.handle), g, C.c_void_p(sv), 0, C.c_size_t(2**31 - 2), 1, 0, C.byref(res), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    for f in ("cugraph_paths_result_get_vertices", "cugraph_paths_result_get_distances"):
        getattr(L, f).restype = C.c_void_p
        getattr(L, f).argtypes = [C.c_void_p]
    verts = _view_to_np(L, L.cugraph_paths_result_get_vertices(res))
    dist = _view_to_np(L, L.cugraph_paths_result_get_distances(res))
    dist_of = dict(zip(verts.tolist(), dist.tolist()))
    imax = np.iinfo(np.int32).max
    unreached = [v for v in verts.tolist() if dist_of[v] == imax]
    assert unreached, "the test graph should have several components"
    r = np.random.default_rng(1)
    dests = np.concatenate([r.choice(verts, 200), np.array(unreached[:5] + [source])]).astype(np.int32)
    dv = L.cugraph_type_erased_device_array_view_create(dests.ctypes.data, dests.size, INT32)
    L.cugraph_extract_paths.argtypes = [C.c_void_p] * 7
    L.cugraph_extract_paths_result_get_max_path_length.restype = C.c_size_t
    L.cugraph_extract_paths_result_get_max_path_length.argtypes = [C.c_void_p]
    L.cugraph_extract_paths_result_get_paths.restype = C.c_void_p
    L.cugraph_extract_paths_result_get_paths.argtypes = [C.c_void_p]
    L.cugraph_extract_paths_result_free.argtypes = [C.c_void_p]
    out = C.c_void_p()
    code = L.cugraph_extract_paths(C.c_void_p(L.handle), g, C.c_void_p(sv), res, C.c_void_p(dv), C.byref(out), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    length = int(L.cugraph_extract_paths_result_get_max_path_length(out))
    paths = _view_to_np(L, L.cugraph_extract_paths_result_get_paths(out)).reshape(dests.size, length)
    reached_d = [dist_of[int(t)] for t in dests if dist_of[int(t)] != imax]
    assert length == 1 + max(reached_d)
    edges = set(zip(s.tolist(), d.tolist()))
    for row, t in zip(paths, dests.tolist()):
        dt = dist_of[t]
        if dt == imax:
            assert (row == -1).all()
            continue
        assert row[0] == source and row[dt] == t and (row[dt + 1:] == -1).all()
        assert all((int(a), int(b)) in edges for a, b in zip(row[:dt], row[1:dt + 1]))
        assert [dist_of[int(x)] for x in row[:dt + 1]] == list(range(dt + 1))
    L.cugraph_extract_paths_result_free(out)
    L.cugraph_paths_result_free(res)
    L.cugraph_graph_free(g)


def test_sssp_single_cta_rounds_emulated(emu, monkeypatch, capfd):  # noqa: F811
    """small near queues are relaxed round after round inside one CTA (k_sssp_small_rounds): same distances, valid predecessors,
    and the path is really taken"""
    s, d = symmetric_edges(6_000, 40_000, seed=11)
    r = np.random.default_rng(2)
    wh = r.random(s.size // 2).astype(np.float32)
    w = np.concatenate([wh, wh])
    g = create_sym_graph(emu, s, d, w)
    ids, ss, dd = dense_ids(s, d)
    source = int(ids[1])
    ref_d, _ = oracle.sssp(ss, dd, w, ids.size, int(np.searchsorted(ids, source)), use_float=True)
    calls = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("CUGRAPH_B200_SSSP_SMALL_ROUNDS", mode)
        monkeypatch.setenv("CUGRAPH_B200_SSSP_TRACE", "1")
        emu.emu_reload_tuning(C.c_void_p(emu.handle))
        capfd.readouterr()
        verts, dist, pred = _sssp_dist(emu, g, source)
        trace = capfd.readouterr().err
        got = np.zeros(ids.size, dtype=np.float32)
        got[np.searchsorted(ids, verts)] = dist
        assert (got == ref_d.astype(np.float32)).all(), mode
        last = [ln for ln in trace.splitlines() if ln.startswith("sssp window")][-1]
        calls[mode] = int(last.split("single-CTA calls")[1].split(")")[0])
    assert calls["1"] > 0 and calls["0"] == 0, calls
    monkeypatch.delenv("CUGRAPH_B200_SSSP_SMALL_ROUNDS")
    monkeypatch.delenv("CUGRAPH_B200_SSSP_TRACE")
    emu.emu_reload_tuning(C.c_void_p(emu.handle))
    emu.cugraph_graph_free(g)


def test_smoke_equivalent_emulated(emu):  # noqa: F811
    """the sequence of __graft_entry__.smoke() (RMAT-12, vertices_array with isolated vertices, PageRank + BFS + SSSP)"""
    from oracle.rmat import rmat_edgelist
    L = emu
    scale = 12
    s, d = rmat_edgelist(scale, 16 << scale, seed=3)
    s, d = np.ascontiguousarray(s, np.int32), np.ascontiguousarray(d, np.int32)
    V = 1 << scale
    verts_all = np.arange(V, dtype=np.int32)

    def mk(src, dst, w, sym, transposed):
        vs = [L.cugraph_type_erased_device_array_view_create(a.ctypes.data, a.size, t) if a is not None else None
              for a, t in ((verts_all, INT32), (src, INT32), (dst, INT32), (w, FLOAT32))]
        g, err = C.c_void_p(), C.c_void_p()
        code = L.cugraph_graph_create_with_times_sg(C.c_void_p(L.handle), C.byref(Props(int(sym), 1)), C.c_void_p(vs[0]),
                                                    C.c_void_p(vs[1]), C.c_void_p(vs[2]), C.c_void_p(vs[3]) if vs[3] else None,
                                                    None, None, None, None, int(transposed), 1, 0, 0, 0, 0, C.byref(g), C.byref(err))
        assert code == 0, L.cugraph_error_message(err)
        return g

    g = mk(s, d, None, False, True)
    verts, pr, it = run_pagerank(L, g, 0.85, 0.0, 20)
    ref, _, _ = oracle.pagerank(s, d, V, None, alpha=0.85, epsilon=0.0, max_iterations=20)
    got = np.zeros(V)
    got[verts] = pr
    assert verts.size == V
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-12)
    L.cugraph_graph_free(g)
    s2, d2 = np.concatenate([s, d]), np.concatenate([d, s])
    wh = np.random.default_rng(0).random(s.shape[0]).astype(np.float32)
    w2 = np.concatenate([wh, wh])
    g2 = mk(s2, d2, w2, True, False)
    src = int(s[0])
    sarr = np.array([src], dtype=np.int32)
    sv = L.cugraph_type_erased_device_array_view_create(sarr.ctypes.data, 1, INT32)
    res, err = C.c_void_p(), C.c_void_p()
    assert L.cugraph_bfs(C.c_void_p(L.handle), g2, C.c_void_p(sv), 1, C.c_size_t(2**31 - 2), 1, 0, C.byref(res), C.byref(err)) == 0
    bv, dist, pred = _paths(L, res)
    rd, _ = oracle.bfs(s2, d2, V, [src])
    gd = np.zeros(V, dtype=np.int32)
    gd[bv] = dist
    assert np.array_equal(gd, rd)
    L.cugraph_sssp.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    res = C.c_void_p()
    assert L.cugraph_sssp(C.c_void_p(L.handle), g2, src, float("inf"), 1, 0, C.byref(res), C.byref(err)) == 0
    sv2, sd, sp = _paths(L, res)
    rs, _ = oracle.sssp(s2, d2, w2, V, src)
    gs = np.zeros(V)
    gs[sv2] = sd
    assert np.array_equal(gs, rs)
    L.cugraph_graph_free(g2)
