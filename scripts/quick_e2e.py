"""Breakdown of the e2e step of bench.py: H2D, graph creation, PageRank (first call builds the blocked
layout and the out-weights), D2H.  python scripts/quick_e2e.py [scale] [reps]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cugraph_b200 import pylibcugraph as plc
from cugraph_b200.generators import rmat_edgelist

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
src, dst = rmat_edgelist(scale, 16 << scale, seed=0)
E = src.numel()
h_src = torch.empty(E, dtype=torch.int32).pin_memory(); h_src.copy_(src)
h_dst = torch.empty(E, dtype=torch.int32).pin_memory(); h_dst.copy_(dst)
del src, dst
h = plc.ResourceHandle()
props = plc.GraphProperties(is_symmetric=False, is_multigraph=True)
def now():
    torch.cuda.synchronize(); return time.perf_counter()
for r in range(reps):
    t0 = now()
    s = h_src.cuda(non_blocking=True); d = h_dst.cuda(non_blocking=True)
    t1 = now()
    g = plc.SGGraph(h, props, s, d, store_transposed=True, renumber=True)
    t2 = now()
    vv, pp, _ = plc.pagerank(h, g, None, None, None, None, 0.85, 0.0, 100, False, fail_on_nonconvergence=False)
    t3 = now()
    a, b = vv.cpu(), pp.cpu()
    t4 = now()
    del g, s, d, vv, pp
    t5 = now()
    print(f"rep {r}: h2d {1e3*(t1-t0):.1f}  create {1e3*(t2-t1):.1f}  pagerank {1e3*(t3-t2):.1f}  d2h {1e3*(t4-t3):.1f}  free {1e3*(t5-t4):.1f}  total {1e3*(t5-t0):.1f} ms"
          f"  -> {E*100/(t4-t0)/1e6:.0f} MTEPS", flush=True)
