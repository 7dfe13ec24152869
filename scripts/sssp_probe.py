import os, sys, time
import torch
sys.path.insert(0, ".")
from cugraph_b200 import pylibcugraph as plc
from cugraph_b200.generators import rmat_edgelist
scale = 24
src, dst = rmat_edgelist(scale, 16 << scale, seed=0)
s2 = torch.cat([src, dst]); d2 = torch.cat([dst, src]); del src, dst
g = torch.Generator(device="cuda"); g.manual_seed(2)
w = torch.rand(s2.numel() // 2, device="cuda", generator=g); w2 = torch.cat([w, w])
h = plc.ResourceHandle()
G = plc.SGGraph(h, plc.GraphProperties(is_symmetric=True, is_multigraph=True), s2, d2, weight_array=w2, renumber=True)
del s2, d2, w2, w
plc.sssp(h, G, 3807250, float("inf"), True, False)
os.environ["CUGRAPH_B200_SSSP_TRACE"] = "1"
torch.cuda.synchronize(); t0 = time.time()
plc.sssp(h, G, 3807250, float("inf"), True, False)
torch.cuda.synchronize(); print(f"sssp traced: {(time.time()-t0)*1e3:.2f} ms", file=sys.stderr)
os.environ.pop("CUGRAPH_B200_SSSP_TRACE")
torch.cuda.synchronize(); t0 = time.time()
plc.sssp(h, G, 3807250, float("inf"), False, False)
torch.cuda.synchronize(); print(f"sssp no-pred: {(time.time()-t0)*1e3:.2f} ms", file=sys.stderr)
