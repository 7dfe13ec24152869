"""Localise a parity failure of an experimental sweep variant: one sweep as configured by the CUGRAPH_B200_* environment
against the plain reference sweep, worst row per class.   CUGRAPH_B200_HOT_X=1 python scripts/debug_variant.py [scale]"""
import ctypes as C
import os
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402
from cugraph_b200 import _capi  # noqa: E402
from cugraph_b200 import pylibcugraph as plc  # noqa: E402
from cugraph_b200.generators import rmat_edgelist  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
os.environ.setdefault("CUGRAPH_B200_SWEEP_MIN_EDGES", "0")
src, dst = rmat_edgelist(scale, 16 << scale, seed=0)
h = plc.ResourceHandle()
g = plc.SGGraph(h, plc.GraphProperties(is_multigraph=True), src, dst, store_transposed=True, renumber=True)
L = _capi.lib()
out = (C.c_double * 8)()
err = C.c_void_p()
f = L.cugraph_b200_debug_compare_sweeps
f.restype = C.c_int
f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
_capi.check(f(h.ptr, g.ptr, C.cast(out, C.c_void_p), C.byref(err)), err, "cugraph_b200_debug_compare_sweeps")
knobs = {k: v for k, v in os.environ.items() if k.startswith("CUGRAPH_B200_")}
print("switches:", knobs)
for k, name in enumerate(("degree >= 32 rows (blocked kernel)", "degree < 32 rows (low kernel)")):
    rel, row, deg, bad = out[4 * k:4 * k + 4]
    print(f"{name:36s} max rel diff {rel:.3e} at row {int(row)} (degree {int(deg)}); rows above 1e-5: {int(bad)}")
