"""Fuzz the emulated library (optionally the AddressSanitizer build) against the oracle: small random multigraphs, random
storage order / renumbering / id width / weight type, PageRank + BFS + SSSP through the C ABI.
    python emu/fuzz.py [seconds] [lib.so]
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python emu/fuzz.py 120 /tmp/libcugraph_c_emu_asan.so"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "emu"))
import oracle  # noqa: E402

INT32, INT64, FLOAT32, FLOAT64 = 2, 3, 8, 9
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
if len(sys.argv) > 2:
    path = sys.argv[2]
else:
    import build_emu
    path = build_emu.build()
L = C.CDLL(path)
L.cugraph_create_resource_handle.restype = C.c_void_p
L.cugraph_create_resource_handle.argtypes = [C.c_void_p]
L.cugraph_type_erased_device_array_view_create.restype = C.c_void_p
L.cugraph_type_erased_device_array_view_create.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
L.cugraph_type_erased_device_array_view_free.argtypes = [C.c_void_p]
for f in ("size", "type", "pointer"):
    fn = getattr(L, f"cugraph_type_erased_device_array_view_{f}")
    fn.restype = {"size": C.c_size_t, "type": C.c_int, "pointer": C.c_void_p}[f]
    fn.argtypes = [C.c_void_p]
L.cugraph_error_message.restype = C.c_char_p
L.cugraph_error_message.argtypes = [C.c_void_p]
L.cugraph_graph_free.argtypes = [C.c_void_p]
for f in ("cugraph_centrality_result_get_vertices", "cugraph_centrality_result_get_values", "cugraph_paths_result_get_vertices",
          "cugraph_paths_result_get_distances", "cugraph_paths_result_get_predecessors"):
    getattr(L, f).restype = C.c_void_p
    getattr(L, f).argtypes = [C.c_void_p]
L.cugraph_centrality_result_free.argtypes = [C.c_void_p]
L.cugraph_paths_result_free.argtypes = [C.c_void_p]
L.cugraph_sssp.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
H = C.c_void_p(L.cugraph_create_resource_handle(None))


class Props(C.Structure):
    _fields_ = [("is_symmetric", C.c_int), ("is_multigraph", C.c_int)]


def view(a):
    if a is None:
        return None
    t = {np.dtype(np.int32): INT32, np.dtype(np.int64): INT64, np.dtype(np.float32): FLOAT32, np.dtype(np.float64): FLOAT64}[a.dtype]
    return C.c_void_p(L.cugraph_type_erased_device_array_view_create(a.ctypes.data, a.size, t))


def to_np(v):
    n = L.cugraph_type_erased_device_array_view_size(v)
    t = L.cugraph_type_erased_device_array_view_type(v)
    dt = {2: np.int32, 3: np.int64, 8: np.float32, 9: np.float64}[t]
    p = L.cugraph_type_erased_device_array_view_pointer(v)
    out = np.ctypeslib.as_array(C.cast(p, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy() if n else np.zeros(0, dt)
    L.cugraph_type_erased_device_array_view_free(C.c_void_p(v))
    return out


def graph(src, dst, w, symmetric, store_transposed, renumber):
    vs, vd, vw = view(src), view(dst), view(w)
    g, err = C.c_void_p(), C.c_void_p()
    code = L.cugraph_graph_create_with_times_sg(H, C.byref(Props(int(symmetric), 1)), None, vs, vd, vw, None, None, None, None,
                                                int(store_transposed), int(renumber), 0, 0, 0, 0, C.byref(g), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    for v in (vs, vd, vw):
        if v:
            L.cugraph_type_erased_device_array_view_free(v)
    return g


def check(cond, what, ctx):
    if not cond:
        print("MISMATCH:", what, ctx, flush=True)
        sys.exit(1)


r = np.random.default_rng(int(os.environ.get("FUZZ_SEED", "0")))
t_end = time.time() + budget
n_cases = 0
while time.time() < t_end:
    if os.environ.get("FUZZ_BIG"):   # several column blocks, many work units: slow under emulation, few cases per minute
        V = int(r.choice([30_000, 60_000, 120_000]))
        E = int(r.integers(2 * V, 6 * V))
    else:
        V = int(r.choice([1, 2, 3, 7, 40, 300, 3000]))
        E = int(r.integers(0, 12 * V + 2))
    idt = r.choice([np.int32, np.int64])
    wt = r.choice([None, np.float32, np.float64])
    renumber = bool(r.integers(0, 2))
    transposed = bool(r.integers(0, 2))
    hot = bool(r.integers(0, 2))
    os.environ["CUGRAPH_B200_SWEEP_MIN_EDGES"] = "0" if hot else "1000000000"
    os.environ["CUGRAPH_B200_SWEEP_BANK_ORDER"] = str(r.choice(["0", "1"]))
    L.emu_reload_tuning(H)   # the knobs are read per handle
    skew = float(r.choice([1.0, 2.5])) if not os.environ.get("FUZZ_BIG") else float(r.choice([2.0, 3.0]))
    src = np.minimum((V * r.random(E) ** skew).astype(np.int64), V - 1).astype(idt)
    dst = np.minimum((V * r.random(E) ** skew).astype(np.int64), V - 1).astype(idt)
    if not renumber and E:
        src[0] = V - 1          # renumber=false: the vertex count is max id + 1
    w = None if wt is None else (r.random(E) + 0.05).astype(wt)
    ctx = dict(V=V, E=E, idt=idt.__name__, wt=None if wt is None else wt.__name__, renumber=renumber, transposed=transposed,
               env={k: v for k, v in os.environ.items() if k.startswith("CUGRAPH_B200")})
    # ---- PageRank on the directed multigraph
    if E > 0:
        g = graph(src, dst, w, False, transposed, renumber)
        res, err = C.c_void_p(), C.c_void_p()
        code = L.cugraph_pagerank_allow_nonconvergence(H, g, None, None, None, None, C.c_double(0.85), C.c_double(0.0),
                                                       C.c_size_t(10), 0, C.byref(res), C.byref(err))
        check(code == 0, L.cugraph_error_message(err) if code else "", ctx)
        verts = to_np(L.cugraph_centrality_result_get_vertices(res))
        pr = to_np(L.cugraph_centrality_result_get_values(res))
        L.cugraph_centrality_result_free(res)
        L.cugraph_graph_free(g)
        if renumber:
            ids, inv = np.unique(np.concatenate([src, dst]), return_inverse=True)
            s, d = inv[:E].astype(np.int32), inv[E:].astype(np.int32)
        else:
            ids = np.arange(int(max(src.max(), dst.max())) + 1)
            s, d = src.astype(np.int32), dst.astype(np.int32)
        ref, _, _ = oracle.pagerank(s, d, ids.size, None if w is None else w.astype(np.float64), alpha=0.85, epsilon=0.0,
                                    max_iterations=10)
        got = np.full(ids.size, np.nan)
        got[np.searchsorted(ids, verts)] = pr
        tol = 1e-5 if wt is not np.float64 else 1e-9
        check(np.allclose(got, ref, rtol=tol, atol=1e-12), f"pagerank max rel {np.nanmax(np.abs(got - ref) / np.maximum(ref, 1e-300)):.2e}", ctx)
    # ---- BFS / SSSP on the symmetrised graph
    if E > 0:
        s2, d2 = np.concatenate([src, dst]), np.concatenate([dst, src])
        w2 = None if w is None else np.concatenate([w, w])
        g = graph(s2, d2, w2, True, False, renumber)
        if renumber:
            ids, inv = np.unique(np.concatenate([s2, d2]), return_inverse=True)
            ss, dd = inv[:2 * E].astype(np.int32), inv[2 * E:].astype(np.int32)
        else:
            ids = np.arange(int(max(s2.max(), d2.max())) + 1)
            ss, dd = s2.astype(np.int32), d2.astype(np.int32)
        source = ids[int(r.integers(0, ids.size))] if renumber else idt(src[int(r.integers(0, E))])
        sidx = int(np.searchsorted(ids, source))
        sarr = np.array([source], dtype=idt)
        sv = view(sarr)
        res, err = C.c_void_p(), C.c_void_p()
        code = L.cugraph_bfs(H, g, sv, int(r.integers(0, 2)), C.c_size_t(2**31 - 2), 1, 0, C.byref(res), C.byref(err))
        check(code == 0, L.cugraph_error_message(err) if code else "", ctx)
        L.cugraph_type_erased_device_array_view_free(sv)
        verts = to_np(L.cugraph_paths_result_get_vertices(res))
        dist = to_np(L.cugraph_paths_result_get_distances(res))
        to_np(L.cugraph_paths_result_get_predecessors(res))
        L.cugraph_paths_result_free(res)
        ref_d, _ = oracle.bfs(ss, dd, ids.size, [sidx])
        got = np.zeros(ids.size, dtype=np.int64)
        got[np.searchsorted(ids, verts)] = dist
        big = np.iinfo(idt).max
        reach = got != big
        check((got[reach] == ref_d[reach]).all() and reach.sum() == (ref_d < np.iinfo(ref_d.dtype).max).sum() if ref_d.dtype.kind == "i"
              else True, "bfs distances", ctx)
        if w2 is not None:
            res, err = C.c_void_p(), C.c_void_p()
            code = L.cugraph_sssp(H, g, int(source), float("inf"), 1, 0, C.byref(res), C.byref(err))
            check(code == 0, L.cugraph_error_message(err) if code else "", ctx)
            verts = to_np(L.cugraph_paths_result_get_vertices(res))
            dist = to_np(L.cugraph_paths_result_get_distances(res))
            to_np(L.cugraph_paths_result_get_predecessors(res))
            L.cugraph_paths_result_free(res)
            ref_d, _ = oracle.sssp(ss, dd, w2, ids.size, sidx, use_float=(wt is np.float32))
            got = np.zeros(ids.size, dtype=dist.dtype)
            got[np.searchsorted(ids, verts)] = dist
            check((got == ref_d.astype(dist.dtype)).all(), "sssp distances", ctx)
        L.cugraph_graph_free(g)
    n_cases += 1
print(f"{n_cases} random cases, no mismatch", flush=True)
