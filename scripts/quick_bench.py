"""Development probe (not the contract bench): RMAT graph -> pull SpMV timing + 100-iteration PageRank."""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, ".")
from cugraph_b200 import _capi  # noqa: E402
from cugraph_b200 import pylibcugraph as plc  # noqa: E402
from cugraph_b200.generators import rmat_edgelist  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
t0 = time.time()
src, dst = rmat_edgelist(scale, 16 << scale, seed=0)
torch.cuda.synchronize()
print(f"rmat gen {time.time()-t0:.2f}s  E={src.numel()}", flush=True)
h = plc.ResourceHandle()
t0 = time.time()
g = plc.SGGraph(h, plc.GraphProperties(is_multigraph=True), src, dst, store_transposed=True, renumber=True)
torch.cuda.synchronize()
print(f"graph create {time.time()-t0:.2f}s", flush=True)
del src, dst
L = _capi.lib()
ms, by, err = C.c_double(), C.c_double(), C.c_void_p()
for rep in range(2):
    code = L.cugraph_b200_time_pull_spmv(h.ptr, g.ptr, 20, C.byref(ms), C.byref(by), C.byref(err))
    _capi.check(code, err, "time_pull_spmv")
    E = 16 << scale
    print(f"pull sweep: {ms.value:.4f} ms  alg bytes {by.value/1e9:.3f} GB  -> {by.value/ms.value/1e6:.1f} GB/s  "
          f"{E/ms.value/1e3:.1f} MTEPS", flush=True)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.time()
    v, p, conv = plc.pagerank(h, g, None, None, None, None, 0.85, 0.0, iters, False, fail_on_nonconvergence=False)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"pagerank {iters} it: {dt*1e3:.1f} ms  {dt*1e3/iters:.4f} ms/it  {E*iters/dt/1e6:.0f} MTEPS  sum={p.double().sum().item():.6f}", flush=True)
print("launches", h.launch_count())
