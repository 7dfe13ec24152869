"""Piece-layout statistics of an RMAT graph on the CPU (emulated staging, CUGRAPH_B200_BUILD_TRACE output): pieces per class
and per range of blocks, slot counts and bytes, for the default and the narrow layout, or for the switch sets given.
    python emu/layout_stats.py [scale] ["CUGRAPH_B200_HOT_NARROW=1,CUGRAPH_B200_HOT_MIN_DEGREE=8" ...]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "emu"))
import build_emu  # noqa: E402
from oracle.rmat import rmat_edgelist  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
L = C.CDLL(build_emu.build())
L.cugraph_create_resource_handle.restype = C.c_void_p
L.cugraph_create_resource_handle.argtypes = [C.c_void_p]
L.cugraph_type_erased_device_array_view_create.restype = C.c_void_p
L.cugraph_type_erased_device_array_view_create.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
L.cugraph_error_message.restype = C.c_char_p
L.cugraph_error_message.argtypes = [C.c_void_p]
L.emu_hot_layout.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
L.emu_reset_layouts.argtypes = [C.c_void_p]
H = C.c_void_p(L.cugraph_create_resource_handle(None))


class Props(C.Structure):
    _fields_ = [("is_symmetric", C.c_int), ("is_multigraph", C.c_int)]


t0 = time.time()
src, dst = rmat_edgelist(scale, 16 << scale, seed=0)
src, dst = np.ascontiguousarray(src, np.int32), np.ascontiguousarray(dst, np.int32)
print(f"rmat scale {scale}: {src.size} edges in {time.time() - t0:.1f} s", flush=True)
vs = C.c_void_p(L.cugraph_type_erased_device_array_view_create(src.ctypes.data, src.size, 2))
vd = C.c_void_p(L.cugraph_type_erased_device_array_view_create(dst.ctypes.data, dst.size, 2))
g, err = C.c_void_p(), C.c_void_p()
t0 = time.time()
code = L.cugraph_graph_create_with_times_sg(H, C.byref(Props(0, 1)), None, vs, vd, None, None, None, None, None, 1, 1, 0, 0, 0, 0,
                                            C.byref(g), C.byref(err))
assert code == 0, L.cugraph_error_message(err)
print(f"emulated staging {time.time() - t0:.1f} s", flush=True)
# ---- the degree < 32 rows: rows / entries per degree, and how many of their sources fall into the first column blocks
L.emu_graph_primary.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
ints8 = (C.c_int64 * 8)()
seg = (C.c_int32 * 8)()
ptrs5 = (C.c_void_p * 5)()
L.emu_graph_primary(g, ints8, seg, ptrs5)
n_rows, nnz = int(ints8[0]), int(ints8[1])
off = np.ctypeslib.as_array(C.cast(ptrs5[0], C.POINTER(C.c_int32)), shape=(n_rows + 1,))
idx = np.ctypeslib.as_array(C.cast(ptrs5[1], C.POINTER(C.c_int32)), shape=(nnz,))
n_hi = int(seg[0])
deg = np.diff(off)
low_idx = idx[off[n_hi]:]
print(f"rows {n_rows}, degree>=32 rows {n_hi} with {int(off[n_hi])} entries; degree<32 rows hold {low_idx.size} entries", flush=True)
hist = np.bincount(deg[n_hi:], minlength=32)
print("degree<32 rows per degree 0..31:", hist.tolist(), flush=True)
W = 49088
for k in (1, 2, 4, 16, 64):
    print(f"  sources of degree<32 rows inside the first {k} column block(s): {100.0 * (low_idx < k * W).mean():.1f} %", flush=True)
hi_idx = idx[:off[n_hi]]
for k in (1, 2, 4, 16, 64):
    print(f"  sources of degree>=32 rows inside the first {k} column block(s): {100.0 * (hi_idx < k * W).mean():.1f} %", flush=True)
os.environ["CUGRAPH_B200_SWEEP_MIN_EDGES"] = "0"
os.environ["CUGRAPH_B200_BUILD_TRACE"] = "1"
configs = sys.argv[2:] or ["CUGRAPH_B200_HOT_NARROW=0", "CUGRAPH_B200_HOT_NARROW=1"]
for cfg in configs:
    for kv in cfg.split(","):
        k, v = kv.split("=")
        os.environ[k] = v
    narrow = cfg
    L.emu_reset_layouts(g)
    ints = (C.c_int64 * 12)()
    ptrs = (C.c_void_p * 10)()
    t0 = time.time()
    sys.stderr.write(f"---- {cfg}\n")
    sys.stderr.flush()
    rc = L.emu_hot_layout(H, g, ints, ptrs)
    print(f"narrow={narrow}: rc={rc} W={ints[0]} B={ints[1]} n_hi={ints[2]} nnz_hi={ints[3]} hot slots={ints[4]} slots={ints[5]} "
          f"subs={ints[6]} units={ints[7]} ({time.time() - t0:.1f} s)", flush=True)
