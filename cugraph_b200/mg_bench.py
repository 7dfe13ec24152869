"""bench.py's N>1 arm: 2D-partitioned PageRank, one process per GPU (launched under torchrun)."""
from __future__ import annotations

import json
import os
import time

import torch
import torch.distributed as dist


def run_mg_pagerank(args, metric, alpha, iters, ClockSampler, peaks):
    from cugraph_b200 import mg
    from cugraph_b200.generators import rmat_edgelist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29517")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    groups = mg.make_groups()
    # weak scaling: 2^28 edge draws per GPU (= the N=1 workload); N=8 is BASELINE's scale-27 configuration
    scale = args.scale if args.scale else 24 + max(0, (world - 1).bit_length())
    E_total = 16 << scale
    E_local = E_total // world
    src, dst = rmat_edgelist(scale, E_local, seed=1000 + rank)
    h_src = h_dst = None
    if world <= 8:
        h_src = torch.empty(E_local, dtype=torch.int32).pin_memory()
        h_dst = torch.empty(E_local, dtype=torch.int32).pin_memory()
        h_src.copy_(src)
        h_dst.copy_(dst)
    G = mg.MGGraph(src, dst, None, groups)
    del src, dst
    torch.cuda.empty_cache()

    def step():
        return G.pagerank(alpha, 0.0, iters)

    def timed(fn, n):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            out = fn()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item()), out

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    l0 = G.handle.launch_count()
    if rank == 0:
        sampler.start()
    wall, (verts, pr, _, _) = timed(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = G.handle.launch_count() - l0
    value = E_total * iters * args.steps / wall / 1e6
    # mass check: PageRank sums to 1 over all ranks
    mass = pr.double().sum().reshape(1)
    dist.all_reduce(mass)

    # roofline of the local sweep (no communication): CUDA events on the stream the kernels run on
    import ctypes as C
    from cugraph_b200.pylibcugraph.utils import View
    xg = torch.full((G.x_elems,), 1.0 / G.part.n_global, dtype=G.dtype, device="cuda")
    yp = torch.zeros(G.span, dtype=G.dtype, device="cuda")
    vx, vy, err = View(xg), View(yp), C.c_void_p()
    for _ in range(3):
        G.lib.cugraph_b200_block_pull_sweep(G.handle.ptr, G.block, vx.ptr, vy.ptr, alpha, C.byref(err))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nsw = 20
    e0.record()
    for _ in range(nsw):
        G.lib.cugraph_b200_block_pull_sweep(G.handle.ptr, G.block, vx.ptr, vy.ptr, alpha, C.byref(err))
    e1.record()
    torch.cuda.synchronize()
    ms_sweep = e0.elapsed_time(e1) / nsw
    alg_bytes = G.num_edges_local * 4.0 + (G.n_rows + 1) * 4.0 + G.n_cols * 4.0 + G.n_rows * 4.0
    peak, peak_src = peaks()
    ach = alg_bytes / (ms_sweep * 1e-3) / 1e9
    t = torch.tensor([ach], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    roofline = {"bound": "hbm", "achieved": float(t.item()), "peak": peak, "unit": "GB/s", "frac": float(t.item()) / peak,
                "traffic": None, "peak_source": peak_src, "kernel": "local block pull sweep (min over ranks)",
                "ms_per_sweep": ms_sweep, "algorithmic_bytes_per_sweep": alg_bytes}

    # e2e: pinned host edge list -> H2D -> 2D partition + block build -> 100 iterations -> D2H
    e2e = None
    if h_src is not None:
        del G
        torch.cuda.empty_cache()

        def e2e_step():
            s = h_src.cuda(non_blocking=True)
            d = h_dst.cuda(non_blocking=True)
            g2 = mg.MGGraph(s, d, None, groups)
            v, p, _, _ = g2.pagerank(alpha, 0.0, iters)
            return v.cpu(), p.cpu()

        e2e_wall, (v, p) = timed(e2e_step, 1)
        e2e = {"value": E_total * iters / e2e_wall / 1e6, "unit": "MTEPS", "h2d_bytes_per_step": 2 * E_total * 4,
               "d2h_bytes_per_step": int(v.numel()) * 8 * world, "steps": 1, "ms_per_step": e2e_wall * 1e3,
               "includes": "pinned H2D of the edge list, 2D partition + block staging, 100 iterations, D2H of results"}
    if rank == 0:
        R, Cc = groups.R, groups.C
        out = {"metric": metric, "value": value, "unit": "MTEPS", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": f"pagerank_rmat{scale}_ef16_100it_2d{R}x{Cc}", "scale": scale, "edge_factor": 16,
                          "num_edges": E_total, "edges_per_gpu": E_local, "alpha": alpha, "iterations": iters,
                          "partition": f"2D {R}x{Cc} (all-gather group {R}, reduce-scatter group {Cc})",
                          "mass": float(mass.item()),
                          "mg_split": os.environ.get("CUGRAPH_B200_MG_SPLIT", "0") == "1",
                          "l2": "inputs per sweep exceed the 126 MB L2; no explicit flush"},
               "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": None}
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
