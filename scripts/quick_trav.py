"""Development probe: BFS / SSSP on a symmetrised RMAT graph, Graph500-style TEPS."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from cugraph_b200 import pylibcugraph as plc  # noqa: E402
from cugraph_b200.generators import rmat_edgelist  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 8
src, dst = rmat_edgelist(scale, 16 << scale, seed=0)
s2 = torch.cat([src, dst])
d2 = torch.cat([dst, src])
del src, dst
g = torch.Generator(device="cuda")
g.manual_seed(2)
w = torch.rand(s2.numel() // 2, device="cuda", generator=g)
w2 = torch.cat([w, w])
h = plc.ResourceHandle()
t0 = time.time()
G = plc.SGGraph(h, plc.GraphProperties(is_symmetric=True, is_multigraph=True), s2, d2, weight_array=w2,
                store_transposed=False, renumber=True)
torch.cuda.synchronize()
print(f"graph create {time.time()-t0:.2f}s  E_sym={s2.numel()}", flush=True)
deg = torch.bincount(s2.long(), minlength=1 << scale)
cand = torch.nonzero(deg > 0).flatten()
torch.manual_seed(1)
sources = cand[torch.randperm(cand.numel(), device="cuda")[:nsrc]].to(torch.int32)
del s2, d2, w2, w
for name in ("bfs", "sssp"):
    teps = []
    for i in range(nsrc + 1):
        s = sources[i % nsrc:i % nsrc + 1].contiguous()
        torch.cuda.synchronize()
        t0 = time.time()
        if name == "bfs":
            dist, pred, verts = plc.bfs(h, G, s, True, 0, True, False)
            reached = dist != 2**31 - 1
        else:
            verts, dist, pred = plc.sssp(h, G, int(s.item()), float("inf"), True, False)
            reached = dist < 3e38
        torch.cuda.synchronize()
        dt = time.time() - t0
        ne = int(deg[verts.long()][reached].sum().item()) // 2
        if i > 0:
            teps.append(ne / dt)
        print(f"{name} src={int(s.item())} {dt*1e3:.2f} ms  reached={int(reached.sum())} edges={ne}  {ne/dt/1e6:.0f} MTEPS", flush=True)
    teps = np.array(teps)
    print(f"{name}: harmonic mean {len(teps)/np.sum(1/teps)/1e6:.0f} MTEPS, mean {teps.mean()/1e6:.0f} MTEPS", flush=True)
