"""Development probe: one RMAT edge list, many graph builds under different env knobs, interleaved over
several rounds (clock / power drift shows up as a difference between rounds, not between configurations).
python scripts/sweep_knobs.py <scale> <rounds> "K=V,K2=V2" "K=V3" ...      ("-" = defaults)"""
import ctypes as C, os, sys
sys.path.insert(0, ".")
import torch
from cugraph_b200 import _capi
from cugraph_b200 import pylibcugraph as plc
from cugraph_b200.generators import rmat_edgelist

scale, rounds = int(sys.argv[1]), int(sys.argv[2])
cfgs = sys.argv[3:] or ["-"]
src, dst = rmat_edgelist(scale, 16 << scale, seed=0)
h = plc.ResourceHandle()
L = _capi.lib()
E = src.numel()

def run(cfg):
    env = dict(kv.split("=") for kv in cfg.split(",")) if cfg != "-" else {}
    for k, v in env.items():
        os.environ[k] = v
    g = plc.SGGraph(h, plc.GraphProperties(is_multigraph=True), src, dst, store_transposed=True, renumber=True)
    ms, by, err = C.c_double(), C.c_double(), C.c_void_p()
    best = 1e9
    for _ in range(3):
        _capi.check(L.cugraph_b200_time_pull_spmv(h.ptr, g.ptr, 20, C.byref(ms), C.byref(by), C.byref(err)), err, "t")
        best = min(best, ms.value)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    plc.pagerank(h, g, None, None, None, None, 0.85, 0.0, 100, False, fail_on_nonconvergence=False)
    torch.cuda.synchronize(); e0.record()
    plc.pagerank(h, g, None, None, None, None, 0.85, 0.0, 100, False, fail_on_nonconvergence=False)
    e1.record(); torch.cuda.synchronize()
    pr = e0.elapsed_time(e1)
    for k in env:
        del os.environ[k]
    del g
    return best, by.value, pr

res = {c: [] for c in cfgs}
for r in range(rounds):
    for c in cfgs:
        ms, by, pr = run(c)
        res[c].append((ms, pr))
        print(f"round {r} {c:48s} sweep {ms:.4f} ms  {by/ms/1e6:7.1f} GB/s   pagerank100 {pr:6.2f} ms  {E*100/pr/1e3:8.0f} MTEPS", flush=True)
for c in cfgs:
    print(f"best {c:50s} sweep {min(x[0] for x in res[c]):.4f}  pagerank100 {min(x[1] for x in res[c]):.2f} ms")
