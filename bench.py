#!/usr/bin/env python
"""bench.py — the measurement contract for the PageRank hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): PageRank on RMAT scale-24 edge-factor-16 (Graph500 a,b,c,
multi-edges and self-loops kept, scrambled ids, unweighted, int32 ids / float32 scores), alpha 0.85,
epsilon 0 (never converges), 100 iterations, graph stored transposed.  N>1: weak scaling with 2^28 edge
draws per GPU (scale 24 + log2 N; N = 8 is BASELINE's scale-27 configuration), 2D edge partition.

A STEP is one call of `cugraph_pagerank_allow_nonconvergence` (100 iterations) through the C-ABI.
  value  = MTEPS = E * iterations * steps / time, graph already resident in HBM (graph creation is
           staging, excluded exactly as the reference's own harness does, pagerank_test.cpp:221-236)
  e2e    = the same metric for the whole reference-facing call sequence with HOST buffers inside the
           timed region: pinned edge list -> H2D -> cugraph_graph_create_with_times_sg -> pagerank ->
           D2H of (vertices, scores)
  roofline = the pull-SpMV sweep (kernels k_spmv_blocked + k_spmv_blocked_finish + k_spmv_low = one
           per_v_transform_reduce_incoming_e) timed alone with CUDA events on the handle's stream;
           algorithmic bytes per sweep = 4E + 4(V+1) + 4V + 4V; traffic = DRAM bytes of
           the three kernels from the ncu launch list (profiles/spmv_traffic.json)
  cpu_baseline = oracle port (oracle/oracle.c, OpenMP) on a bounded sample, host cores stated
  timing = CUDA events recorded on the handle's stream (the stream the library launches on) around the K calls;
           the wall-clock time of the same region is reported next to it
  side   = informational extras measured in separate processes under timeouts (scripts/bench_side.py): BFS / SSSP
           TEPS on the symmetrised RMAT-24 graph (BASELINE.json configs[2], [3]) and the experimental sweep
           variants; CUGRAPH_B200_BENCH_SIDE=0 skips them
Synthetic data, random seed 0.  Inputs (1.2 GB per sweep) exceed the 126 MB L2, so no explicit L2 flush
is needed between timed iterations (stated in config.l2).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALPHA, ITERS = 0.85, 100
METRIC = "MTEPS (million traversed edges/sec) PageRank RMAT-24 ef-16, 100 iterations"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def _cpu_baseline(sample_scale=21, target_s=12.0):
    """Oracle port on the host cores: float32 pull-SpMV sweeps over a smaller RMAT CSC."""
    import numpy as np
    import oracle
    from oracle.rmat import rmat_edgelist
    src, dst = rmat_edgelist(sample_scale, 16 << sample_scale, seed=0)
    V = 1 << sample_scale
    csc = oracle.coo_to_csx(dst, src, V)
    x = np.full(V, 1.0 / V, dtype=np.float32)
    E = src.shape[0]
    t0 = time.perf_counter()
    oracle.spmv_f32(csc, x, ALPHA, 0.0)
    one = time.perf_counter() - t0
    n = max(1, min(50, int(target_s / max(one, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(n):
        x = oracle.spmv_f32(csc, x, ALPHA, 0.15 / V)
    dt = time.perf_counter() - t0
    out = {"value": E * n / dt / 1e6, "unit": "MTEPS", "cores": oracle.num_threads(), "kind": "port",
           "sample": f"{n} float32 pull-SpMV sweeps (oracle_spmv_f32, OpenMP) over RMAT scale-{sample_scale} ef-16 CSC"}
    out["networkx"] = _networkx_baseline(min(sample_scale, 16))
    return out


def _networkx_baseline(scale):
    """NetworkX (BASELINE.json's named CPU baseline): nx.pagerank, 100 power iterations (tol = 0 never converges; the
    PowerIterationFailedConvergence after max_iter marks the end), on a small RMAT sample — building a NetworkX graph at
    benchmark scale is prohibitive.  Graph construction is not timed."""
    try:
        import networkx as nx
        from oracle.rmat import rmat_edgelist
        src, dst = rmat_edgelist(scale, 16 << scale, seed=0)
        G = nx.MultiDiGraph()
        G.add_edges_from(zip(src.tolist(), dst.tolist()))
        t0 = time.perf_counter()
        try:
            nx.pagerank(G, alpha=ALPHA, tol=0.0, max_iter=ITERS)
        except nx.PowerIterationFailedConvergence:
            pass
        dt = time.perf_counter() - t0
        return {"value": src.shape[0] * ITERS / dt / 1e6, "unit": "MTEPS", "cores": 1, "version": nx.__version__,
                "sample": f"nx.pagerank, {ITERS} iterations, RMAT scale-{scale} ef-16 MultiDiGraph (SciPy-backed, single thread)"}
    except Exception as ex:
        return {"value": None, "error": f"{type(ex).__name__}: {ex}"[:200]}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port
    (libcugraph itself is not buildable here).  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import numpy as np
    import oracle
    from oracle.rmat import rmat_edgelist
    scale = 20
    src, dst = rmat_edgelist(scale, 16 << scale, seed=0)
    V, E = 1 << scale, src.shape[0]
    csc = oracle.coo_to_csx(dst, src, V)

    def step():
        oracle.pagerank(src, dst, V, None, alpha=ALPHA, epsilon=0.0, max_iterations=ITERS, csc=csc)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = E * ITERS * args.steps / dt / 1e6
    sample = f"PageRank {ITERS} iterations on RMAT scale-{scale} ef-16 per step (oracle_pagerank, fp64, OpenMP)"
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "MTEPS", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "pagerank_rmat24_ef16_100it", "sample_scale": scale},
           "cpu_baseline": {"value": val, "unit": "MTEPS", "cores": oracle.num_threads(), "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "MTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def run_single(args):
    import torch
    from cugraph_b200 import _capi
    from cugraph_b200 import pylibcugraph as plc
    from cugraph_b200.generators import rmat_edgelist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(0)
    scale = args.scale
    L = _capi.lib()
    src, dst = rmat_edgelist(scale, 16 << scale, seed=0)
    E = src.numel()
    h = plc.ResourceHandle()
    props = plc.GraphProperties(is_symmetric=False, is_multigraph=True)
    G = plc.SGGraph(h, props, src, dst, store_transposed=True, renumber=True)
    # pinned host copy of the edge list for the e2e arm
    h_src = torch.empty(E, dtype=torch.int32).pin_memory()
    h_dst = torch.empty(E, dtype=torch.int32).pin_memory()
    h_src.copy_(src)
    h_dst.copy_(dst)
    del src, dst

    def step():
        return plc.pagerank(h, G, None, None, None, None, ALPHA, 0.0, ITERS, False, fail_on_nonconvergence=False)

    for _ in range(args.warmup):
        v, p, _ = step()
    nv = v.numel()
    sampler = ClockSampler(0)
    torch.cuda.synchronize()
    sampler.start()
    l0 = h.launch_count()
    # CUDA events on the stream the library launches on (the handle's own stream, not torch's current stream)
    hstream = torch.cuda.ExternalStream(int(L.cugraph_b200_handle_stream(h.ptr) or 0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(hstream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()  # synchronous on return (the C-ABI syncs the handle's stream)
    e1.record(hstream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    launches = h.launch_count() - l0
    clocks = sampler.stop()
    dev_s = e0.elapsed_time(e1) * 1e-3  # device time between the two events; wall is the host's view of the same region
    # the calls are synchronous, so the two clocks must agree; if the events saw something else (wrong stream), fall back
    events_ok = 0.5 * wall <= dev_s <= 1.05 * wall
    timed_s = dev_s if events_ok else wall
    ms_step = timed_s / args.steps * 1e3
    value = E * ITERS * args.steps / timed_s / 1e6

    # roofline: the pull sweep alone, CUDA events on the handle's stream inside the library
    ms, by, err = C.c_double(), C.c_double(), C.c_void_p()
    code = L.cugraph_b200_time_pull_spmv(h.ptr, G.ptr, 50, C.byref(ms), C.byref(by), C.byref(err))
    _capi.check(code, err, "cugraph_b200_time_pull_spmv")
    peak, peak_src = _peaks()
    achieved = by.value / (ms.value * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_sweep")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "pull sweep: k_spmv_blocked+k_spmv_blocked_finish+k_spmv_low",
                "ms_per_sweep": ms.value, "algorithmic_bytes_per_sweep": by.value,
                "sweep_mteps": E / (ms.value * 1e-3) / 1e6}
    # the same ratio for a whole PageRank iteration (B_iter = B_spmv + 4 V-sized streams of the vertex pass)
    b_iter = by.value + 16.0 * nv
    ms_iter = ms_step / ITERS
    roofline["iteration"] = {"algorithmic_bytes": b_iter, "ms": ms_iter, "achieved": b_iter / (ms_iter * 1e-3) / 1e9,
                             "frac": b_iter / (ms_iter * 1e-3) / 1e9 / peak}

    # the device-resident numbers go to stderr right away: should anything below take the process down, the log has them
    sys.stderr.write("[bench provisional] " + json.dumps({"value": value, "unit": "MTEPS", "ms_per_step": ms_step,
                                                          "gpu_launches": launches, "roofline": roofline}) + "\n")
    sys.stderr.flush()

    # e2e: host edge list -> H2D -> graph create -> pagerank -> D2H
    del G
    torch.cuda.empty_cache()
    e2e_steps = max(1, min(args.steps, 3))
    # pinned host buffers for the result (a pageable .cpu() costs 20-30 ms for 134 MB)
    h_v = torch.empty(nv, dtype=torch.int32).pin_memory()
    h_p = torch.empty(nv, dtype=torch.float32).pin_memory()

    def e2e_step():
        s = h_src.cuda(non_blocking=True)
        d = h_dst.cuda(non_blocking=True)
        g = plc.SGGraph(h, props, s, d, store_transposed=True, renumber=True)  # waits for the two copies first
        vv, pp, _ = plc.pagerank(h, g, None, None, None, None, ALPHA, 0.0, ITERS, False, fail_on_nonconvergence=False)
        assert vv.numel() == nv, f"e2e graph has {vv.numel()} vertices, the resident one {nv}"
        h_v.copy_(vv, non_blocking=True)
        h_p.copy_(pp, non_blocking=True)
        torch.cuda.synchronize()
        return h_v, h_p

    try:
        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_wall = time.perf_counter() - t0
        e2e = {"value": E * ITERS * e2e_steps / e2e_wall / 1e6, "unit": "MTEPS", "h2d_bytes_per_step": 2 * E * 4,
               "d2h_bytes_per_step": nv * 8, "steps": e2e_steps, "ms_per_step": e2e_wall / e2e_steps * 1e3,
               "includes": "pinned H2D of edge list, graph staging, 100 iterations, D2H of vertices+scores into pinned buffers"}
    except Exception as ex:  # keep the device-resident measurement even if the host-buffer arm fails
        e2e = {"value": None, "unit": "MTEPS", "h2d_bytes_per_step": 2 * E * 4, "d2h_bytes_per_step": nv * 8,
               "error": f"{type(ex).__name__}: {ex}"[:300]}

    del h_src, h_dst, h_v, h_p
    torch.cuda.empty_cache()
    side = None
    if os.environ.get("CUGRAPH_B200_BENCH_SIDE", "1") != "0":
        try:
            side = _side_measurements(scale)
        except Exception as ex:  # informational extras must never cost the main line
            side = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    try:
        cpu = _cpu_baseline(sample_scale=args.cpu_sample_scale)
    except Exception as ex:  # the GPU measurements above must still be reported
        cpu = {"value": None, "unit": "MTEPS", "cores": None, "kind": "port", "error": f"{type(ex).__name__}: {ex}"[:300]}
    out = {"metric": METRIC, "value": value, "unit": "MTEPS", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": f"pagerank_rmat{scale}_ef16_100it", "scale": scale, "edge_factor": 16,
                      "num_vertices": nv, "num_edges": E, "alpha": ALPHA, "iterations": ITERS, "vertex_type": "int32",
                      "l2": "inputs (1.2 GB/sweep) exceed the 126 MB L2; no explicit flush"},
           "timing": {"device_ms_per_step": dev_s / args.steps * 1e3, "wall_ms_per_step": wall / args.steps * 1e3,
                      "value_from": "device events" if events_ok else "wall clock (events disagreed)",
                      "how": "CUDA events recorded on the handle's stream around the K synchronous C-ABI calls"},
           "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu}
    if side is not None:
        out["side"] = side
    print(json.dumps(out), flush=True)


# Side measurements (informational; the headline keys above never depend on them).  Each runs in its own process under
# a timeout: BFS / SSSP of BASELINE.json configs[2] and [3] on the default path, then the experimental sweep variants
# (off by default, parity-checked against the plain sweep in the same process before they are timed).
SIDE_VARIANTS = ["-", "CUGRAPH_B200_HOT_X=1", "CUGRAPH_B200_HOT_BANK_ORDER=1", "CUGRAPH_B200_HOT_X=1,CUGRAPH_B200_HOT_NARROW=1",
                 "CUGRAPH_B200_LOW_ELL=1",
                 "CUGRAPH_B200_LOW_ELL=2", "CUGRAPH_B200_HOT_X=1,CUGRAPH_B200_HOT_NARROW=1,CUGRAPH_B200_LOW_ELL=2",
                 "CUGRAPH_B200_HOT_X=1,CUGRAPH_B200_HOT_NARROW=1,CUGRAPH_B200_HOT_BANK_ORDER=1,CUGRAPH_B200_LOW_ELL=2",
                 "CUGRAPH_B200_HOT_X=1,CUGRAPH_B200_HOT_NARROW=1,CUGRAPH_B200_HOT_MIN_DEGREE=8",
                 "CUGRAPH_B200_HOT_X=1,CUGRAPH_B200_HOT_NARROW=1,CUGRAPH_B200_HOT_MIN_DEGREE=1", "CUGRAPH_B200_LOW_ASYNC=1"]


def _run_side(argv, timeout_s):
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "bench_side.py")] + [str(a) for a in argv]
    t0 = time.perf_counter()
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, cwd=ROOT)
    except subprocess.TimeoutExpired:
        return {"error": f"timeout after {timeout_s} s"}
    except Exception as ex:
        return {"error": f"{type(ex).__name__}: {ex}"[:300]}
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if r.returncode != 0 or not lines:
        return {"error": f"exit code {r.returncode}", "stderr_tail": r.stderr[-400:]}
    try:
        res = json.loads(lines[-1])
    except Exception as ex:
        return {"error": f"unparsable output: {ex}"[:200]}
    res["process_s"] = time.perf_counter() - t0
    return res


def _side_measurements(scale, budget_s=None):
    budget_s = float(os.environ.get("CUGRAPH_B200_BENCH_SIDE_BUDGET_S", "130")) if budget_s is None else budget_s
    t0 = time.perf_counter()
    side = {"traversal": _run_side(["traversal", scale, 16, 4], 120), "variants": []}
    for cfg in SIDE_VARIANTS:
        if time.perf_counter() - t0 > budget_s:
            side["variants"].append({"config": cfg, "skipped": f"side budget of {budget_s:.0f} s spent"})
            continue
        res = _run_side(["variant", scale, cfg], 50)
        res.setdefault("config", cfg)
        side["variants"].append(res)
    side["seconds"] = time.perf_counter() - t0
    side["note"] = ("informational: default-path BFS/SSSP (Graph500 TEPS, random sources) and experimental sweep variants "
                    "(each parity-checked against the plain sweep, then timed); the headline keys use the default path only")
    return side


def run_multi(args):
    from cugraph_b200.mg_bench import run_mg_pagerank
    run_mg_pagerank(args, METRIC, ALPHA, ITERS, ClockSampler, _peaks)


def _protect_stdout():
    """Libraries (NCCL's version banner, torchrun children) write to fd 1; the contract is ONE JSON line on
    stdout.  Route fd 1 to stderr for the run and keep the real stdout for the final line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--scale", type=int, default=None, help="override RMAT scale (development only)")
    ap.add_argument("--cpu-sample-scale", type=int, default=21, help="RMAT scale of the cpu_baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        return run_multi(args)  # weak scaling: scale 24 + log2(N); N = 8 is BASELINE's scale-27 configuration
    if args.scale is None:
        args.scale = 24
    return run_single(args)


if __name__ == "__main__":
    main()
