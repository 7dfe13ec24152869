"""Development probe: BFS level trace for slow sources + SSSP delta sweep."""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from cugraph_b200 import pylibcugraph as plc  # noqa: E402
from cugraph_b200.generators import rmat_edgelist  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
src, dst = rmat_edgelist(scale, 16 << scale, seed=0)
s2 = torch.cat([src, dst]); d2 = torch.cat([dst, src])
del src, dst
g = torch.Generator(device="cuda"); g.manual_seed(2)
w = torch.rand(s2.numel() // 2, device="cuda", generator=g)
w2 = torch.cat([w, w])
h = plc.ResourceHandle()
G = plc.SGGraph(h, plc.GraphProperties(is_symmetric=True, is_multigraph=True), s2, d2, weight_array=w2, renumber=True)
del s2, d2, w2, w
for s in (10775680, 8063756, 3807250):
    st = torch.tensor([s], dtype=torch.int32, device="cuda")
    plc.bfs(h, G, st, True, 0, True, False)
    torch.cuda.synchronize(); t0 = time.time()
    plc.bfs(h, G, st, True, 0, True, False)
    torch.cuda.synchronize(); print(f"bfs {s}: {(time.time()-t0)*1e3:.2f} ms", file=sys.stderr, flush=True)
os.environ.pop("CUGRAPH_B200_BFS_TRACE", None)
for sc in ("1", "0.5", "0.25", "0.1", "0.05", "0.02"):
    os.environ["CUGRAPH_B200_SSSP_DELTA_SCALE"] = sc
    plc.sssp(h, G, 3807250, float("inf"), True, False)
    torch.cuda.synchronize(); t0 = time.time()
    v, d, p = plc.sssp(h, G, 3807250, float("inf"), True, False)
    torch.cuda.synchronize()
    print(f"sssp delta_scale {sc}: {(time.time()-t0)*1e3:.2f} ms  max finite dist {d[d < 3e38].max().item():.4f}", file=sys.stderr, flush=True)
