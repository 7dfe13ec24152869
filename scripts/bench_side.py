"""Side measurements of bench.py, each run in its own process (a crash or a hang of one of them cannot lose the
main bench line).  Prints ONE JSON line.

    python scripts/bench_side.py traversal <scale> <n_bfs_sources> <n_sssp_sources>
        BFS (direction-optimising) and SSSP on the symmetrised RMAT graph, Graph500-style TEPS per source
        (undirected edges of the source's component / time of the C-ABI call), 1 warm-up source,
        harmonic + arithmetic mean; the first result of each algorithm is checked with size-independent
        properties (BFS: dist[pred[v]] + 1 == dist[v]; SSSP: dist[pred[v]] <= dist[v], dist[source] == 0).

    python scripts/bench_side.py variant <scale> <K=V,K2=V2 | ->
        one configuration of the CUGRAPH_B200_* switches: parity of one pull sweep against the plain reference
        sweep (cugraph_b200_debug_compare_sweeps), then the sweep time (cugraph_b200_time_pull_spmv, CUDA events on
        the handle's stream) and one PageRank call of 100 iterations.
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _harmonic(xs):
    return len(xs) / sum(1.0 / x for x in xs) if xs else None


def traversal(scale, n_bfs, n_sssp):
    import torch
    from cugraph_b200 import pylibcugraph as plc
    from cugraph_b200.generators import rmat_edgelist
    torch.cuda.set_device(0)
    V = 1 << scale
    src, dst = rmat_edgelist(scale, 16 << scale, seed=0)
    s2, d2 = torch.cat([src, dst]), torch.cat([dst, src])
    del src, dst
    g = torch.Generator(device="cuda")
    g.manual_seed(2)
    w = torch.rand(s2.numel() // 2, device="cuda", generator=g)
    w2 = torch.cat([w, w])
    del w
    h = plc.ResourceHandle()
    t0 = time.perf_counter()
    G = plc.SGGraph(h, plc.GraphProperties(is_symmetric=True, is_multigraph=True), s2, d2, weight_array=w2,
                    store_transposed=False, renumber=True)
    torch.cuda.synchronize()
    create_s = time.perf_counter() - t0
    deg = torch.bincount(s2.long(), minlength=V)
    e_sym = int(s2.numel())
    del s2, d2, w2
    cand = torch.nonzero(deg > 0).flatten()
    torch.manual_seed(1)
    n_src = max(n_bfs, n_sssp)
    sources = cand[torch.randperm(cand.numel(), device="cuda")[:n_src + 1]].to(torch.int32)
    out = {"graph": {"scale": scale, "symmetrised_edges": e_sym, "create_s": create_s}}

    def run(name, n, env=None, check=True):
        """n timed sources after one warm-up source; env = schedule knobs (they never change results)"""
        env = env or {}
        os.environ.update(env)
        teps, ms, checked = [], [], None
        l0 = h.launch_count()
        try:
            for i in range(n + 1):  # source 0 is the warm-up
                s = sources[i:i + 1].contiguous()
                s_host = int(s.item())
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                if name == "bfs":
                    dist, pred, verts = plc.bfs(h, G, s, True, 0, True, False)
                else:
                    verts, dist, pred = plc.sssp(h, G, s_host, float("inf"), True, False)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                reached = (dist != 2**31 - 1) if name == "bfs" else (dist < 3e38)
                ne = int(deg[verts.long()][reached].sum().item()) // 2
                if i == 0:
                    if check:  # property check on the full-size result (external ids)
                        d_ext = torch.empty(V, dtype=dist.dtype, device="cuda")
                        d_ext[verts.long()] = dist
                        has_pred = pred >= 0
                        dp = d_ext[pred[has_pred].long()]
                        dv = dist[has_pred]
                        ok_tree = bool(((dp + 1 == dv) if name == "bfs" else (dp <= dv)).all().item())
                        ok_src = bool((d_ext[s_host] == 0).item())
                        ok_cnt = int(has_pred.sum().item()) == int(reached.sum().item()) - 1
                        checked = {"tree_property": ok_tree, "source_distance_zero": ok_src,
                                   "every_reached_vertex_but_the_source_has_a_predecessor": ok_cnt,
                                   "reached": int(reached.sum().item())}
                else:
                    teps.append(ne / dt)
                    ms.append(dt * 1e3)
        finally:
            for k in env:
                os.environ.pop(k, None)
        res = {"sources": n, "harmonic_mean_mteps": _harmonic(teps) / 1e6 if teps else None,
               "mean_mteps": sum(teps) / len(teps) / 1e6 if teps else None,
               "mean_ms": sum(ms) / len(ms) if ms else None, "min_ms": min(ms) if ms else None,
               "max_ms": max(ms) if ms else None, "launches_per_source": (h.launch_count() - l0) / (n + 1)}
        if check:
            res["check"] = checked
            res["timing"] = "wall clock around the synchronous C-ABI call, torch.cuda.synchronize() on both sides"
        if env:
            res["env"] = env
        return res

    out["bfs"] = run("bfs", n_bfs)
    out["sssp"] = run("sssp", n_sssp)
    # schedule knobs A/B on the same graph and sources (results are identical by construction; only time differs)
    ab = []
    for name, n, env in (("bfs", n_bfs, {"CUGRAPH_B200_BFS_ALPHA": "40"}), ("bfs", n_bfs, {"CUGRAPH_B200_BFS_ALPHA": "120"}),
                         ("sssp", min(n_sssp, 2), {"CUGRAPH_B200_SSSP_ADAPTIVE": "0"}),
                         ("sssp", min(n_sssp, 2), {"CUGRAPH_B200_SSSP_SPLIT_ROUNDS": "2"})):
        try:
            r = run(name, n, env, check=False)
            r["algorithm"] = name
            ab.append(r)
        except Exception as ex:
            ab.append({"algorithm": name, "env": env, "error": f"{type(ex).__name__}: {ex}"[:200]})
    out["schedule_ab"] = ab
    print(json.dumps(out), flush=True)


def variant(scale, cfg):
    env = dict(kv.split("=") for kv in cfg.split(",")) if cfg != "-" else {}
    os.environ.update(env)
    import torch
    from cugraph_b200 import _capi
    from cugraph_b200 import pylibcugraph as plc
    from cugraph_b200.generators import rmat_edgelist
    torch.cuda.set_device(0)
    L = _capi.lib()
    src, dst = rmat_edgelist(scale, 16 << scale, seed=0)
    E = src.numel()
    h = plc.ResourceHandle()
    g = plc.SGGraph(h, plc.GraphProperties(is_multigraph=True), src, dst, store_transposed=True, renumber=True)
    del src, dst
    out = {"config": cfg}
    # 1. parity of one sweep against the plain reference sweep
    cmp8 = (C.c_double * 8)()
    err = C.c_void_p()
    f = L.cugraph_b200_debug_compare_sweeps
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    _capi.check(f(h.ptr, g.ptr, C.cast(cmp8, C.c_void_p), C.byref(err)), err, "cugraph_b200_debug_compare_sweeps")
    out["parity"] = {"max_rel_diff_degree_ge32": cmp8[0], "rows_above_1e-5_degree_ge32": int(cmp8[3]),
                     "max_rel_diff_degree_lt32": cmp8[4], "rows_above_1e-5_degree_lt32": int(cmp8[7])}
    out["parity_ok"] = int(cmp8[3]) == 0 and int(cmp8[7]) == 0
    # 2. sweep time (best of 3 x 20 sweeps) and one PageRank call
    ms, by = C.c_double(), C.c_double()
    best = None
    for _ in range(3):
        _capi.check(L.cugraph_b200_time_pull_spmv(h.ptr, g.ptr, 20, C.byref(ms), C.byref(by), C.byref(err)), err,
                    "cugraph_b200_time_pull_spmv")
        best = ms.value if best is None else min(best, ms.value)
    out["sweep_ms"] = best
    out["sweep_gbs"] = by.value / (best * 1e-3) / 1e9
    plc.pagerank(h, g, None, None, None, None, 0.85, 0.0, 100, False, fail_on_nonconvergence=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v, p, _ = plc.pagerank(h, g, None, None, None, None, 0.85, 0.0, 100, False, fail_on_nonconvergence=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["pagerank100_ms"] = dt * 1e3
    out["pagerank_mteps"] = E * 100 / dt / 1e6
    out["pagerank_mass"] = float(p.double().sum().item())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "traversal":
        traversal(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    elif mode == "variant":
        variant(int(sys.argv[2]), sys.argv[3])
    else:
        raise SystemExit(f"unknown mode {mode}")
