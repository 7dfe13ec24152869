"""Multi-GPU PageRank and BFS: 2D edge partition over one process per GPU (torch.distributed, NCCL on NVLink 5).

What the reference does: P = R x C GPUs, vertex -> GPU by hash
(cpp/include/cugraph/utilities/graph_partition_utils.cuh:30-43, 101-128), every GPU holds the edge
block (row block of destinations) x (column block of sources); per PageRank iteration the scaled
ranks are all-gathered inside the column group (update_edge_src_property, grouped ncclBroadcast,
update_edge_src_dst_property.cuh:550-579), the local pull sweep runs, and the partial sums are reduced
to their owners inside the row group (grouped ncclReduce, per_v_transform_reduce_e.cuh:3389-3407).

Here: the same partition, but the exchange is ONE all_gather_into_tensor + ONE reduce_scatter_tensor
per iteration on equal-sized (padded) vertex partitions, the dangling / convergence scalars travel in
one 2-element all_reduce and never touch the host unless epsilon > 0, and the local sweep is the
same column-blocked shared-memory kernel as on one GPU (C-ABI: cugraph_b200_block_*).

`partition_edges` (pure torch, device agnostic: exercised on CPU with the gloo backend in
tests/test_mg_partition_cpu.py) builds the blocks; `MGGraph` / `pagerank` need CUDA.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


# ----------------------------------------------------------------------------------------------
# process grid (reference: cpp/tests/utilities/mg_utilities.cpp:49-53, partition_manager.hpp:106-114)
# ----------------------------------------------------------------------------------------------
def grid_shape(world: int):
    """(R, C) of the P = R x C process grid: R = size of the all-gather (column) group, C = size of the reduce (row) group.
    The two factors are the reference's (largest divisor <= sqrt(P) and its cofactor, mg_utilities.cpp:49-53); the LARGER one
    is R here: 2 -> 2x1, 4 -> 2x2, 8 -> 4x2.  A GPU's block has C * maxpart destination rows and R * maxpart source columns.
    The number of pieces (= accumulations) of a block does not depend on the orientation (numpy count on RMAT, scale 20 per
    GPU: 1x2 5.12 M pieces / 2x1 5.20 M; 2x4 6.88 M / 4x2 6.98 M) but its rows do: with R >= C the accumulator array and
    the finish pass are half (N = 2, 8) the size — 70 MB instead of 141 MB of accumulators at RMAT-25 on 2 GPUs, i.e. inside
    the L2 again.  CUGRAPH_B200_MG_GRID=wide restores the reference's orientation (R <= C)."""
    r = int(math.isqrt(world))
    while world % r:
        r -= 1
    small, large = r, world // r
    if os.environ.get("CUGRAPH_B200_MG_GRID", "tall") == "wide":
        return small, large
    return large, small


def vertex_owner(ext: torch.Tensor, world: int) -> torch.Tensor:
    """Balanced vertex -> GPU map (the role of murmurhash3_32(ext) % P in the reference)."""
    x = ext.to(torch.int64)
    x = (x ^ (x >> 30)) * -4658895280553007687   # 0xbf58476d1ce4e5b9 as int64
    x = (x ^ (x >> 27)) * -7723592293110705685   # 0x94d049bb133111eb as int64
    x = x ^ (x >> 31)
    return (x & 0x7FFFFFFFFFFFFFFF) % world


@dataclass
class Groups:
    world: int
    rank: int
    R: int
    C: int
    r: int
    c: int
    row_group: object   # ranks (r, *)  : reduce-scatter of partial y
    col_group: object   # ranks (*, c)  : all-gather of x


def make_groups() -> Groups:
    world, rank = dist.get_world_size(), dist.get_rank()
    R, Cc = grid_shape(world)
    r, c = rank // Cc, rank % Cc
    row_group = col_group = None
    for rr in range(R):
        g = dist.new_group([rr * Cc + cc for cc in range(Cc)])
        if rr == r:
            row_group = g
    for cc in range(Cc):
        g = dist.new_group([rr * Cc + cc for rr in range(R)])
        if cc == c:
            col_group = g
    return Groups(world, rank, R, Cc, r, c, row_group, col_group)


# ----------------------------------------------------------------------------------------------
# collectives that also run on gloo (CPU tests)
# ----------------------------------------------------------------------------------------------
def _is_nccl(group=None):
    return dist.get_backend(group) == "nccl"


def exchange(tensors, dest: torch.Tensor, world: int):
    """all-to-all-v: element i of every tensor goes to rank dest[i]. Returns received tensors."""
    order = torch.argsort(dest, stable=True)
    send_counts = torch.bincount(dest, minlength=world).to(torch.int64)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    out = []
    for t in tensors:
        ts = t[order].contiguous()
        tr = torch.empty(sum(rc), dtype=t.dtype, device=t.device)
        dist.all_to_all_single(tr, ts, output_split_sizes=rc, input_split_sizes=sc)
        out.append(tr)
    return out, order, sc, rc


def all_gather_into(out: torch.Tensor, inp: torch.Tensor, group):
    if _is_nccl(group):
        dist.all_gather_into_tensor(out, inp, group=group)
    else:
        n = dist.get_world_size(group)
        parts = [torch.empty_like(inp) for _ in range(n)]
        dist.all_gather(parts, inp, group=group)
        out.copy_(torch.cat(parts))


def reduce_scatter_into(out: torch.Tensor, inp: torch.Tensor, group, op=None):
    op = op or dist.ReduceOp.SUM
    if _is_nccl(group):
        dist.reduce_scatter_tensor(out, inp, op=op, group=group)
    else:
        tmp = inp.clone()
        dist.all_reduce(tmp, op=op, group=group)
        k = dist.get_rank(group)
        out.copy_(tmp[k * out.numel():(k + 1) * out.numel()])


def _unique_big(tensors, chunk=1 << 29):
    """sorted unique of the concatenation, in pieces (torch.unique is limited to < 2^31 elements)."""
    parts = []
    for t in tensors:
        for i in range(0, max(t.numel(), 1), chunk):
            parts.append(torch.unique(t[i:i + chunk]))
    while len(parts) > 1:
        parts = [torch.unique(torch.cat(parts[i:i + 2])) for i in range(0, len(parts), 2)]
    return parts[0]


# ----------------------------------------------------------------------------------------------
# 2D partition
# ----------------------------------------------------------------------------------------------
@dataclass
class Partition:
    groups: Groups
    rows: torch.Tensor        # int32, local destination slot of every local edge: c_v * maxpart + lid
    cols: torch.Tensor        # int32, local source slot: r_u * maxpart + lid (partition-major: the all-gather output order)
    weights: object           # tensor or None
    vertices: torch.Tensor    # external ids of the vertices this rank owns, in local-id order
    n_local: int
    maxpart: int
    n_global: int


def partition_edges(src: torch.Tensor, dst: torch.Tensor, weights=None, groups: Groups | None = None) -> Partition:
    """Shuffle this rank's share of the edge list into the 2D partition and renumber.
    Edge (u -> v) is stored on GPU (r(owner(v)), c(owner(u)))  (graph_partition_utils.cuh:101-128)."""
    g = groups or make_groups()
    P, Cc = g.world, g.C
    so, do = vertex_owner(src, P), vertex_owner(dst, P)
    target = (do // Cc) * Cc + (so % Cc)
    payload = [src, dst] + ([weights] if weights is not None else [])
    recv, _, _, _ = exchange(payload, target, P)
    src_e, dst_e = recv[0], recv[1]
    w_e = recv[2] if weights is not None else None
    # vertices referenced here -> their owners, together with how often each is a SOURCE here; the owner
    # numbers its vertices by descending global out-degree so that hot sources get the lowest local ids
    # (the column-blocked sweep keeps the lowest column ids in shared memory) — the role of the
    # reference's per-GPU degree-descending renumbering (renumber_edgelist_impl.cuh:732-738)
    u = _unique_big([src_e, dst_e])
    ks = torch.searchsorted(u, src_e)
    cnt = torch.bincount(ks, minlength=u.numel()).to(torch.int64)
    ou = vertex_owner(u, P)
    (recv_ids, recv_cnt), order, sc, rc = exchange([u, cnt], ou, P)
    mine, inv = torch.unique(recv_ids, return_inverse=True)   # sorted external ids owned by this rank
    n_local = int(mine.numel())
    deg = torch.zeros(n_local, dtype=torch.int64, device=src.device).index_add_(0, inv, recv_cnt)
    by_deg = torch.argsort(-deg, stable=True)                  # by_deg[l] = index into `mine` of local id l
    lid_of_mine = torch.empty_like(by_deg)
    lid_of_mine[by_deg] = torch.arange(n_local, device=src.device)
    pos = lid_of_mine[inv]                                     # local id of every requested vertex
    mine = mine[by_deg]                                        # external ids in local-id order
    back = torch.empty(sum(sc), dtype=pos.dtype, device=pos.device)
    dist.all_to_all_single(back, pos.contiguous(), output_split_sizes=sc, input_split_sizes=rc)
    lid = torch.empty_like(back)
    lid[order] = back                                   # lid[k] = local id (at its owner) of u[k]
    t = torch.tensor([n_local, n_local], dtype=torch.int64, device=src.device)
    mx = t[:1].clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    tot = t[1:].clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    maxpart = max(int(mx.item()), 1)
    kd = torch.searchsorted(u, dst_e)
    rows = ((ou[kd] % Cc) * maxpart + lid[kd]).to(torch.int32)
    # columns are partition-major (slot = r_u * maxpart + lid): exactly the layout all_gather_into_tensor produces, so the
    # gathered x is consumed in place (an interleaved order put all hot sources into the first column block but cost a
    # strided 4 * n_cols-byte transpose copy per iteration; every partition's hot sources still lead ITS column range)
    cols = ((ou[ks] // Cc) * maxpart + lid[ks]).to(torch.int32)
    return Partition(g, rows, cols, w_e, mine, n_local, maxpart, int(tot.item()))


# ----------------------------------------------------------------------------------------------
# CUDA side
# ----------------------------------------------------------------------------------------------
def _view(t):
    from cugraph_b200.pylibcugraph.utils import View
    return View(t)


class MGGraph:
    """This rank's edge block of a 2D-partitioned graph (pull orientation: rows = destinations)."""

    def __init__(self, src, dst, weights=None, groups: Groups | None = None, dtype=torch.float32):
        from cugraph_b200 import _capi
        from cugraph_b200.pylibcugraph.resource_handle import ResourceHandle
        assert src.is_cuda or _capi.emulated(), "MGGraph needs CUDA tensors"
        self.lib = _capi.lib()
        self._capi = _capi
        self.dtype = dtype if weights is None else weights.dtype
        self.part = partition_edges(src, dst, weights, groups)
        p = self.part
        g = p.groups
        self.handle = ResourceHandle(stream=torch.cuda.current_stream().cuda_stream)
        self.n_rows, self.n_cols = g.C * p.maxpart, g.R * p.maxpart
        es = 4 if self.dtype == torch.float32 else 8
        # EXPERIMENTAL (CUGRAPH_B200_MG_SPLIT=1): one block per destination partition of the row group, so that the
        # reduction of partition j's partial sums runs under the sweep of partition j+1 (pagerank_split)
        self.split = os.environ.get("CUGRAPH_B200_MG_SPLIT", "0") == "1" and g.C > 1
        self.block, self.blocks = None, []

        def make_block(rows, cols, w, n_rows):
            rv, cv, wv = _view(rows), _view(cols), _view(w)
            blk, err = C.c_void_p(), C.c_void_p()
            code = self.lib.cugraph_b200_block_create(self.handle.ptr, n_rows, self.n_cols, rv.ptr, cv.ptr, wv.ptr,
                                                      C.byref(blk), C.byref(err))
            for v in (rv, cv, wv):
                v.free()
            _capi.check(code, err, "cugraph_b200_block_create")
            return blk.value

        if self.split:
            part_of = torch.div(p.rows, p.maxpart, rounding_mode="floor")
            for j in range(g.C):
                m = part_of == j
                rj = (p.rows[m] - j * p.maxpart).to(p.rows.dtype).contiguous()
                cj = p.cols[m].contiguous()
                wj = p.weights[m].contiguous() if p.weights is not None else None
                self.blocks.append(make_block(rj, cj, wj, p.maxpart))
            self.spans = [int(self.lib.cugraph_b200_block_span(b)) for b in self.blocks]
            self.span = max(self.spans)
        else:
            self.block = make_block(p.rows, p.cols, p.weights, self.n_rows)
            self.span = int(self.lib.cugraph_b200_block_span(self.block))
        self.x_elems = int(self.lib.cugraph_b200_padded_elems(self.span, es))
        # out-weight sums of the owned vertices: partial per column slot, reduce-scattered in the column group
        ones = p.weights.to(torch.float64) if p.weights is not None else torch.ones(p.cols.numel(), dtype=torch.float64, device=src.device)
        partial = torch.zeros(self.n_cols, dtype=torch.float64, device=src.device)
        partial.index_add_(0, p.cols.long(), ones)                          # partition-major already
        ow = torch.empty(p.maxpart, dtype=torch.float64, device=src.device)
        reduce_scatter_into(ow, partial, g.col_group)
        self.out_w = ow.to(self.dtype)
        self.num_edges_local = int(p.rows.numel())
        self.device = src.device
        p.rows = p.cols = p.weights = None  # the block owns its own copy
        torch.cuda.synchronize()

    def __del__(self):
        try:
            if getattr(self, "block", None):
                self.lib.cugraph_b200_block_free(self.block)
                self.block = None
            for b in getattr(self, "blocks", []):
                self.lib.cugraph_b200_block_free(b)
            self.blocks = []
        except Exception:
            pass

    # one PageRank iteration = all-gather(x) -> block sweep -> reduce-scatter(y) -> vertex step -> all-reduce(2 scalars)
    def pagerank(self, alpha=0.85, epsilon=1e-5, max_iterations=100, time_iterations=False):
        if self.split:
            return self.pagerank_split(alpha, epsilon, max_iterations)
        p, g, L, capi = self.part, self.part.groups, self.lib, self._capi
        dev, dt, mp = self.out_w.device, self.dtype, p.maxpart
        pr = torch.zeros(mp, dtype=dt, device=dev)
        pr[:p.n_local] = 1.0 / p.n_global
        x_local = torch.zeros(mp, dtype=dt, device=dev)
        xg = torch.zeros(self.x_elems, dtype=dt, device=dev)
        ypart = torch.zeros(self.span, dtype=dt, device=dev)
        yred = torch.zeros(mp, dtype=dt, device=dev)
        tot = torch.zeros(2, dtype=torch.float64, device=dev)
        part = torch.zeros(2, dtype=torch.float64, device=dev)
        views = {k: _view(v) for k, v in dict(pr=pr, x=x_local, xg=xg, yp=ypart, yr=yred, ow=self.out_w).items()}
        err = C.c_void_p()

        pending = [None]

        def vertex_step(first):
            # the totals of the previous step are needed now: their all-reduce ran under the sweep in between
            if pending[0] is not None:
                pending[0].wait()
                pending[0] = None
            part.zero_()
            code = L.cugraph_b200_pagerank_vertex_step(self.handle.ptr, views["yr"].ptr, views["pr"].ptr, views["ow"].ptr,
                                                       views["x"].ptr, p.n_local, float(alpha), float(p.n_global),
                                                       1 if first else 0, C.c_void_p(tot.data_ptr()),
                                                       C.c_void_p(part.data_ptr()), C.byref(err))
            capi.check(code, err, "cugraph_b200_pagerank_vertex_step")
            pending[0] = dist.all_reduce(part, async_op=True)

        vertex_step(True)
        tot, part = part, tot
        iters, converged = 0, False
        xcols = xg[:self.n_cols]
        for _ in range(int(max_iterations)):
            if g.R == 1:
                xcols[:mp].copy_(x_local)
            else:
                all_gather_into(xcols, x_local, g.col_group)            # partition-major = the block's column order
            code = L.cugraph_b200_block_pull_sweep(self.handle.ptr, self.block, views["xg"].ptr, views["yp"].ptr,
                                                   float(alpha), C.byref(err))
            capi.check(code, err, "cugraph_b200_block_pull_sweep")
            if g.C == 1:
                yred.copy_(ypart[:mp])
            else:
                reduce_scatter_into(yred, ypart[:self.n_rows], g.row_group)
            vertex_step(False)
            tot, part = part, tot
            iters += 1
            if epsilon > 0.0:   # host sync only when a tolerance is requested
                pending[0].wait()
                pending[0] = None
                if float(tot[0].item()) < epsilon:
                    break
        if pending[0] is not None:
            pending[0].wait()
        converged = iters < max_iterations
        for v in views.values():
            v.free()
        return p.vertices, pr[:p.n_local].clone(), iters, converged

    # ------------------------------------------------------------------------------------------
    # multi-GPU BFS.  The reference's MG BFS (bfs_impl.cuh:446-869) moves the frontier through the edge partitions with
    # fill_edge_dst_property broadcasts (fill_edge_src_dst_property.cuh:1368) and an all-to-all-v of the discovered
    # (vertex, predecessor) pairs over the row communicator (transform_reduce_if_v_frontier_outgoing_e_by_dst.cuh:981-1074).
    # Here one level is: all-gather of the owners' frontier flags inside the column group (-> flags over the block's
    # source slots), all-gather of the visited flags inside the row group (-> flags over its destination slots), the
    # block's pull step on the device (cugraph_b200_block_bfs_pull: every unvisited destination looks for a source in the
    # frontier), ONE max-reduce-scatter of the candidate predecessors inside the row group, and the owners' update.
    # Distances are the BFS levels (bit-exact vs single-GPU); a predecessor is any frontier neighbour, as in the reference.
    # ------------------------------------------------------------------------------------------
    def bfs(self, source, depth_limit=-1, compute_predecessors=True):
        """source: external vertex id (the same value on every rank).  Returns (vertices, distances, predecessors) of the
        vertices this rank owns: int32 distances (INT32_MAX = unreachable), predecessors as external ids (-1 = none)."""
        assert self.block is not None, "bfs needs the unsplit block"
        p, g, L, capi = self.part, self.part.groups, self.lib, self._capi
        dev, mp = self.device, p.maxpart
        imax = torch.iinfo(torch.int32).max
        dist_own = torch.full((mp,), imax, dtype=torch.int32, device=dev)
        pred_code = torch.full((mp,), -1, dtype=torch.int64, device=dev)
        visited = torch.zeros(mp, dtype=torch.uint8, device=dev)
        visited[p.n_local:] = 1                                  # padding slots never take part
        frontier = torch.zeros(mp, dtype=torch.uint8, device=dev)
        owner = int(vertex_owner(torch.tensor([int(source)], dtype=torch.int64), g.world)[0])
        found = torch.zeros(1, dtype=torch.int64, device=dev)
        if owner == g.rank:
            hit = (p.vertices == int(source)).nonzero()
            if hit.numel():
                lid = int(hit[0, 0])
                dist_own[lid] = 0
                visited[lid] = 1
                frontier[lid] = 1
                found += 1
        dist.all_reduce(found)
        if int(found.item()) == 0:
            raise ValueError(f"bfs source {source} is not a vertex of the graph")
        f_cols = torch.zeros(self.n_cols, dtype=torch.uint8, device=dev)
        v_rows = torch.zeros(self.n_rows, dtype=torch.uint8, device=dev)
        cand = torch.full((self.n_rows,), -1, dtype=torch.int64, device=dev)
        cand_own = torch.full((mp,), -1, dtype=torch.int64, device=dev)
        views = {k: _view(v) for k, v in dict(f=f_cols, v=v_rows, c=cand).items()}
        err = C.c_void_p()
        level = 0
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        while depth_limit < 0 or level < depth_limit:
            if g.R == 1:
                f_cols.copy_(frontier)
            else:
                all_gather_into(f_cols, frontier, g.col_group)
            if g.C == 1:
                v_rows.copy_(visited)
            else:
                all_gather_into(v_rows, visited, g.row_group)
            code = L.cugraph_b200_block_bfs_pull(self.handle.ptr, self.block, views["f"].ptr, views["v"].ptr, mp, g.C, g.c,
                                                 views["c"].ptr, C.byref(err))
            capi.check(code, err, "cugraph_b200_block_bfs_pull")
            if g.C == 1:
                cand_own.copy_(cand)
            else:
                reduce_scatter_into(cand_own, cand, g.row_group, op=dist.ReduceOp.MAX)
            new = (visited == 0) & (cand_own >= 0)
            level += 1
            dist_own[new] = level
            pred_code[new] = cand_own[new]
            visited |= new.to(torch.uint8)
            frontier = new.to(torch.uint8)
            count.fill_(int(new.sum().item()))
            dist.all_reduce(count)
            if int(count.item()) == 0:
                break
        for v in views.values():
            v.free()
        verts = p.vertices
        d_out = dist_own[:p.n_local].clone()
        if not compute_predecessors:
            return verts, d_out, None
        # predecessor codes (owner rank * maxpart + local id) -> external ids, answered by the owners
        codes = pred_code[:p.n_local]
        has = codes >= 0
        ask = codes[has]
        (req,), order, sc, rc = exchange([ask % mp], torch.div(ask, mp, rounding_mode="floor"), g.world)
        ans = verts[req]
        back = torch.empty(sum(sc), dtype=ans.dtype, device=dev)
        dist.all_to_all_single(back, ans.contiguous(), output_split_sizes=sc, input_split_sizes=rc)
        got = torch.empty_like(back)
        got[order] = back
        pred = torch.full((p.n_local,), -1, dtype=verts.dtype, device=dev)
        pred[has] = got
        return verts, d_out, pred


# EXPERIMENTAL: same iteration, but the block is split by destination partition: sweep j, then an asynchronous
# reduce of its partial sums to member j of the row group while sweep j+1 runs (the reference's per-block
# ncclReduce, per_v_transform_reduce_e.cuh:3389-3407).  Only the last reduce and the all-gather stay exposed.
def _pagerank_split(self, alpha=0.85, epsilon=1e-5, max_iterations=100):
    p, g, L, capi = self.part, self.part.groups, self.lib, self._capi
    dev, dt, mp = self.out_w.device, self.dtype, p.maxpart
    pr = torch.zeros(mp, dtype=dt, device=dev)
    pr[:p.n_local] = 1.0 / p.n_global
    x_local = torch.zeros(mp, dtype=dt, device=dev)
    xg = torch.zeros(self.x_elems, dtype=dt, device=dev)
    ybufs = [torch.zeros(sp, dtype=dt, device=dev) for sp in self.spans]
    ymine = ybufs[g.c][:mp]                      # the reduction for this rank's vertices lands here
    tot = torch.zeros(2, dtype=torch.float64, device=dev)
    part = torch.zeros(2, dtype=torch.float64, device=dev)
    views = {k: _view(v) for k, v in dict(pr=pr, x=x_local, xg=xg, yr=ymine, ow=self.out_w).items()}
    yviews = [_view(y) for y in ybufs]
    err = C.c_void_p()

    def vertex_step(first):
        part.zero_()
        code = L.cugraph_b200_pagerank_vertex_step(self.handle.ptr, views["yr"].ptr, views["pr"].ptr, views["ow"].ptr,
                                                   views["x"].ptr, p.n_local, float(alpha), float(p.n_global),
                                                   1 if first else 0, C.c_void_p(tot.data_ptr()),
                                                   C.c_void_p(part.data_ptr()), C.byref(err))
        capi.check(code, err, "cugraph_b200_pagerank_vertex_step")
        dist.all_reduce(part)

    vertex_step(True)
    tot, part = part, tot
    iters = 0
    for _ in range(int(max_iterations)):
        if g.R == 1:
            xg[:mp].copy_(x_local)
        else:
            all_gather_into(xg[:self.n_cols], x_local, g.col_group)
        works = []
        for j in range(g.C):
            code = L.cugraph_b200_block_pull_sweep(self.handle.ptr, self.blocks[j], views["xg"].ptr, yviews[j].ptr,
                                                   float(alpha), C.byref(err))
            capi.check(code, err, "cugraph_b200_block_pull_sweep")
            works.append(dist.reduce(ybufs[j][:mp], dst=g.r * g.C + j, group=g.row_group, async_op=True))
        for wk in works:
            wk.wait()
        vertex_step(False)
        tot, part = part, tot
        iters += 1
        if epsilon > 0.0 and float(tot[0].item()) < epsilon:
            break
    converged = iters < max_iterations
    for v in list(views.values()) + yviews:
        v.free()
    return p.vertices, pr[:p.n_local].clone(), iters, converged


MGGraph.pagerank_split = _pagerank_split


def bfs(graph: MGGraph, source, depth_limit=-1, compute_predecessors=True):
    """(vertices, distances, predecessors) of the vertices owned by this rank (the MG contract of pylibcugraph.bfs)."""
    return graph.bfs(source, depth_limit, compute_predecessors)


def pagerank(graph: MGGraph, alpha=0.85, epsilon=1e-5, max_iterations=100):
    """Returns (vertices, pageranks, converged) for the vertices owned by this rank
    (the MG contract of pylibcugraph.pagerank: every rank gets its local part)."""
    v, p, it, conv = graph.pagerank(alpha, epsilon, max_iterations)
    return v, p, conv
